import csv, sys, subprocess, collections, re, os, tempfile
rep, lib, ksub = sys.argv[1:4]
iters = float(sys.argv[4])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]; data = rows[hi + 1:]
iE = hdr.index('Instructions Executed')
ex = [int(r[iE] or 0) for r in data]
n = len(ex)
for thr in (0.05, 0.25, 0.5, 0.9, 2, 10, 30):
    k = sum(1 for v in ex if v >= thr * iters)
    print(f"executed >= {thr:5.2f} x per iteration: {k:6d} instr = {k*16/1024:6.1f} kB")
# footprint in 128-B lines touched at least 0.25x per iteration
lines = collections.Counter()
for i, v in enumerate(ex):
    if v >= 0.25 * iters: lines[i // 8] += 1
print("128B lines with an instruction executed >= 0.25x/iter:", len(lines), "=", len(lines) * 128 / 1024, "kB; total lines", (n + 7) // 8)
# map to functions
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(lib)} >/dev/null && nvdisasm -g -hex *.cubin > dis.txt", shell=True, check=True)
dis = open(f"{tmp}/dis.txt").read().splitlines()
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and ksub in l][0]
cur = ('?', 0); seq = []
for l in dis[start + 1:]:
    if l.startswith('//-----'): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): seq.append(cur)
files = {}
def func_of(f, ln):
    if f not in files:
        p = os.path.join('/root/repo/mpc_racing_b200/csrc', f)
        files[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = files[f]
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = re.match(r'\s*(?:template\s*<[^>]*>\s*)?(?:MPCR_D|MPCR_HD|MPCR_NOINLINE|MPCR_MATH)\s+.*?(\w+)\(', L[i])
        if m and not L[i].strip().startswith('//'): return f.split('.')[0][5:] + ':' + m.group(1)
    return f
hot = collections.Counter()
for i in range(min(n, len(seq))):
    if 0.25 * iters <= ex[i] < 10 * iters: hot[func_of(*seq[i])] += 1
print("once-per-iteration code (0.25x..10x) by function:")
for k, v in hot.most_common(25): print(f"  {k:30s} {v:5d} instr {v*16/1024:6.1f} kB")
