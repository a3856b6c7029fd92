import csv, sys, collections
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows[:8]) if '# Samples' in r][0]
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = collections.Counter(); n = 0
    for r in rows[hi+1:]:
        if len(r) <= ix['# Samples'] or not r[ix['# Samples']].isdigit(): continue
        n += int(r[ix['# Samples']])
        for h in cols:
            v = r[ix[h]]
            if v.isdigit(): tot[h] += int(v)
    print(path, n, ' '.join('%s:%.1f%%' % (k.replace('stall_', ''), 100*v/n) for k, v in tot.most_common(10)))
