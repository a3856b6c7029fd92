import csv, re, sys, collections
csvp, disp, fn = sys.argv[1:4]
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
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = ins.match(l)
    if m: seq.append((int(m.group(1), 16), cur, m.group(2)))
rows = list(csv.reader(open(csvp)))
hi = [i for i, r in enumerate(rows[:8]) if '# Samples' in r][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi+1:] if len(r) > ix['# Samples'] and r[ix['# Samples']].isdigit()]
def phase(loc):
    if loc is None: return 'none'
    f, ln = loc
    if f == 'amdf16.cuh':  # (line ranges of the tree this table was last made from: see the commit)
        if ln <= 122: return 'amdf blocks (sad/loads)'
        if ln <= 156: return 'amdf resolve/udiv'
        if ln <= 194: return 'amdf search setup+part store'
        if ln <= 246: return 'amdf pick'
        if ln <= 297: return 'amdf decimate'
        return 'amdf find_pitch glue'
    if f == 'k4_sonic.cu':
        if 124 <= ln <= 138: return 'ensure'
        if 140 <= ln <= 148: return 'advance_out'
        if 151 <= ln <= 194: return 'emit_copy'
        if 195 <= ln <= 280: return 'overlap_add'
        if 577 <= ln <= 682: return 'process()'
        if 699 <= ln <= 849: return 'kernel setup'
        if 850 <= ln <= 1008: return 'event loop'
        return 'k4 other %d' % (ln // 50 * 50)
    if f == 'common.cuh': return 'stage (refill)'
    return f
agg = collections.defaultdict(lambda: [0, 0]); tot = toti = 0
for r, (addr, loc, text) in zip(data, seq):
    s = int(r[ix['# Samples']]); n = int(r[ix['Instructions Executed']] or 0)
    p = phase(loc); agg[p][0] += s; agg[p][1] += n; tot += s; toti += n
iters = float(sys.argv[4]); cyc = float(sys.argv[5])
for p, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(p.ljust(32), "%5.1f%%" % (100 * s / tot), "%7.1f inst/iter" % (n / iters), "%7.0f cyc/iter" % (s / tot * cyc))
