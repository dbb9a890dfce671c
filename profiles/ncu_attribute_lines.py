"""Attribute executed instructions / stall samples of a profiled kernel to source functions.
usage: ncu_attribute_lines.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> [function-to-list]"""
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
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): seq.append(cur)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]; data = rows[hi + 1:]
iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
iT = hdr.index('Thread Instructions Executed') if 'Thread Instructions Executed' in hdr else None
print(len(seq), len(data))
n = min(len(seq), len(data))
files = {}
def lines_of(f):
    if f not in files:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mpc_racing_b200', 'csrc', f)
        files[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return files[f]
def func_of(f, ln):
    L = lines_of(f)
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = re.match(r'\s*(?:template\s*<[^>]*>\s*)?(?:MPCR_D|MPCR_HD|MPCR_NOINLINE|MPCR_MATH)\s+.*?(\w+)\(', L[i])
        if m and not L[i].strip().startswith('//'): return f.split('.')[0][5:] + ':' + m.group(1)
    return f
agg = collections.Counter(); aggs = collections.Counter(); aggt = collections.Counter()
byl = collections.Counter(); byls = collections.Counter()
for k in range(n):
    ex = int(data[k][iE] or 0); sm = int(data[k][iS] or 0)
    th = int(data[k][iT] or 0) if iT is not None else 0
    key = func_of(*seq[k]); agg[key] += ex; aggs[key] += sm; aggt[key] += th
    byl[seq[k]] += ex; byls[seq[k]] += sm
tot = sum(agg.values()); tots = sum(aggs.values()); tott = max(1, sum(aggt.values()))
print('total inst', tot, 'samples', tots, 'thread inst', tott)
for k, v in agg.most_common(40):
    print(f"{k:34s} inst {100*v/tot:6.2f}%  samples {100*aggs[k]/tots:6.2f}%  lanes/inst {aggt[k]/max(1,v):5.1f}")
if len(sys.argv) > 4:
    fn = sys.argv[4]
    for (f, ln), v in sorted(byl.items()):
        if func_of(f, ln).endswith(':' + fn) and (v or byls[(f, ln)]):
            print(f"{ln:5d} {100*v/tot:6.2f}% {100*byls[(f,ln)]/tots:6.2f}%  {lines_of(f)[ln-1][:110]}")
