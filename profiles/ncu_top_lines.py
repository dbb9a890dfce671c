import csv, re, sys, collections, subprocess, os, tempfile
rep, lib, ksub = sys.argv[1:4]
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
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]; data = rows[hi + 1:]
iE = hdr.index("Instructions Executed"); iS = hdr.index(sys.argv[4])
byl = collections.Counter(); byls = collections.Counter()
for k in range(min(len(seq), len(data))):
    byl[seq[k]] += int(data[k][iE] or 0); byls[seq[k]] += int(data[k][iS] or 0)
tot = sum(byl.values()); tots = sum(byls.values())
files = {}
def line(f, ln):
    if f not in files:
        p = os.path.join('/root/repo/mpc_racing_b200/csrc', f)
        files[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = files[f]
    return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ''
for (f, ln), v in byls.most_common(45):
    print(f"{f[:18]:18s}{ln:5d} inst {100*byl[(f,ln)]/tot:5.2f}% samples {100*v/tots:5.2f}%  {line(f, ln)}")
