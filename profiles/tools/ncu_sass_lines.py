"""Map ncu SASS-page samples to CUDA source lines using nvdisasm -g line info.
usage: ncu_sass_lines.py <ncu sass csv> <nvdisasm file> <function substring> [topn]"""
import csv, re, sys, collections
csvp, disp, fn = sys.argv[1:4]; topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- parse nvdisasm: sequence of (line) per instruction for the function
lines = open(disp).read().split('\n')
start = None
for i, l in enumerate(lines):
    if l.startswith('.text.') and fn in l and l.rstrip().endswith(':'):
        start = i; break
cur = None; seq = []
inl = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
ins = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);')
for l in lines[start+1:]:
    if l.startswith('.text.') or l.startswith('//-----'):
        if seq: break
    m = inl.search(l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = ins.match(l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2)))
# --- parse ncu csv
rows = list(csv.reader(open(csvp)))
hi = [i for i, r in enumerate(rows[:8]) if '# Samples' in r][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi+1:] if len(r) > ix['# Samples'] and r[ix['# Samples']].isdigit()]
assert abs(len(data) - len(seq)) < 8, (len(data), len(seq))
agg = collections.defaultdict(lambda: [0, 0]); tot = 0; toti = 0
stall_cols = [h for h in hdr if h.startswith('stall_')]
stall_agg = collections.defaultdict(lambda: collections.Counter())
for r, (addr, loc, text) in zip(data, seq):
    s = int(r[ix['# Samples']]); n = int(r[ix['Instructions Executed']] or 0)
    agg[loc][0] += s; agg[loc][1] += n; tot += s; toti += n
    for h in stall_cols:
        v = r[ix[h]]
        if v.isdigit() and int(v): stall_agg[loc][h] += int(v)
print("total samples", tot, "total warp inst", toti)
src = {}
import os
keyi = 1 if os.environ.get('BY_INST') else 0
for loc, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][keyi])[:topn]:
    if loc is None: print(s, n, None); continue
    f, ln = loc
    if f not in src:
        try: src[f] = open(os.environ.get('SRC_DIR','speedy_b200/csrc/') + f).read().split('\n')
        except Exception: src[f] = []
    text = src[f][ln-1].strip() if 0 < ln <= len(src[f]) else ''
    top_st = ','.join('%s:%d' % (k.replace('stall_', ''), v) for k, v in stall_agg[loc].most_common(3))
    print(str(s).rjust(7), ("%.1f%%" % (100*s/tot)).rjust(6), str(n).rjust(12), (f + ':' + str(ln)).ljust(22), text[:70].ljust(70), top_st)
