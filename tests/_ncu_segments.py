"""Summarise an `ncu --page source --csv --print-source sass` export: per kernel, runs of SASS with equal execution count."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
minM = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
ks = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
seen = set()
for a, b in zip(ks[:-1], ks[1:]):
    body = [r for r in rows[a + 2:b] if len(r) > 8 and r[5].isdigit()]
    tot = sum(int(r[5]) for r in body)
    key = (rows[a][1], tot)
    if key in seen or not body: continue
    seen.add(key)
    print('==', rows[a][1][:60], 'total %.1fM warp-inst, thread/warp %.1f' % (tot / 1e6, sum(int(r[6]) for r in body) / max(1, tot)))
    segs, seg, prev = [], [], None
    for i, r in enumerate(body):
        c = int(r[5])
        if prev is not None and abs(c - prev) > 0.03 * max(c, prev, 1):
            segs.append(seg); seg = []
        seg.append((i, r)); prev = c
    segs.append(seg)
    for seg in segs:
        t = sum(int(r[5]) for _, r in seg)
        if t < minM * 1e6: continue
        th = sum(int(r[6]) for _, r in seg) / t
        ops = collections.Counter(([x for x in r[1].split() if not x.startswith('@')][0].split('.')[0]) for _, r in seg)
        print('  sass %4d-%4d n=%3d count=%8d thr=%4.1f tot=%6.1fM %4.1f%%  %s' % (seg[0][0], seg[-1][0], len(seg), int(seg[0][1][5]), th, t / 1e6, 100 * t / tot, dict(ops.most_common(5))))
