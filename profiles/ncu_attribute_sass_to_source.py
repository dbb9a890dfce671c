# attribute executed instructions / stall samples of the profiled kernel to source lines
import csv, re, sys, collections
dis = open('/tmp/cub/dis.txt').read().splitlines()
kname = '.text._ZN4mpcr12solve_kernelILi0ELb0EEE'
start = [i for i,l in enumerate(dis) if l.startswith(kname)][0]
cur = ('?',0); seq = []   # list of (file,line) per instruction in order
inl = None
for l in dis[start+1:]:
    if l.startswith('//-----') : break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        seq.append(cur)
rows = list(csv.reader(open('/tmp/src.csv')))
hdr = rows[1]; data = rows[2:]
iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
print(len(seq), len(data))
n = min(len(seq), len(data))
by_line = collections.Counter(); by_line_s = collections.Counter()
for k in range(n):
    ex = int(data[k][iE] or 0); sm = int(data[k][iS] or 0)
    by_line[seq[k]] += ex; by_line_s[seq[k]] += sm
tot = sum(by_line.values()); tots = sum(by_line_s.values())
print('total inst', tot, 'samples', tots)
# group by file and line ranges
def region(f, ln):
    if f == 'mpcr_model.cuh':
        if ln < 160: return 'model:f_*'
        if ln < 420: return 'model:lin_*'
        return 'model:cost/poly'
    if f == 'mpcr_ipm.cuh':
        src = R
        for name, lo, hi in src:
            if lo <= ln <= hi: return 'ipm:'+name
        return 'ipm:other'
    return f
# function line ranges in mpcr_ipm.cuh
ipm = open('/root/repo/mpc_racing_b200/csrc/mpcr_ipm.cuh').read().splitlines()
marks = []
for i,l in enumerate(ipm, 1):
    m = re.match(r'\s*MPCR_D\s+.*?(\w+)\(', l) or re.match(r'\s*MPCR_HD\s+.*?(\w+)\(', l)
    if m and not l.strip().startswith('//'): marks.append((m.group(1), i))
R = []
for j,(nm,ln) in enumerate(marks):
    hi = marks[j+1][1]-1 if j+1 < len(marks) else len(ipm)
    R.append((nm, ln, hi))
agg = collections.Counter(); aggs = collections.Counter()
for (f,ln),v in by_line.items(): agg[region(f,ln)] += v
for (f,ln),v in by_line_s.items(): aggs[region(f,ln)] += v
for k,v in agg.most_common(30):
    print(f"{k:32s} inst {100*v/tot:6.2f}%  samples {100*aggs[k]/tots:6.2f}%")
if len(sys.argv) > 1:
    fn = sys.argv[1]
    lo,hi = [(l,h) for n_,l,h in R if n_==fn][0]
    for ln in range(lo,hi+1):
        v = by_line.get(('mpcr_ipm.cuh',ln),0); s_ = by_line_s.get(('mpcr_ipm.cuh',ln),0)
        if v or s_: print(f"{ln:5d} {100*v/tot:6.2f}% {100*s_/tots:6.2f}%  {ipm[ln-1][:100]}")
