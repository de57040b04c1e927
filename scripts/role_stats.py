"""Summarise the per-role cycle counters a build with -DLLC_ROLE_TIMING prints (one line per warp)."""
import collections, re, sys
import numpy as np
for f in sys.argv[1:]:
    rows = [tuple(map(int, re.findall(r'-?\d+', l))) for l in open(f) if l.startswith('cta ')]
    rows = rows[len(rows) // 2:]                     # the last (timed) launch of a two-launch run
    sp = collections.defaultdict(list)
    for c, sm, r, w, t, wid in rows:
        sp[(sm, wid % 4)].append(r)
    tot = np.array([t for c, sm, r, w, t, wid in rows if r == 0])
    print(f, 'warps', len(rows), 'CTA cycles min/mean/max %.0fM %.0fM %.0fM' % (tot.min() / 1e6, tot.mean() / 1e6, tot.max() / 1e6))
    print('  chains per sub-partition', dict(collections.Counter(v.count(0) for v in sp.values())))
    for r in range(max(x[2] for x in rows) + 1):
        w = np.array([x[3] for x in rows if x[2] == r])
        print('  role %d busy cycles mean %.0fM max %.0fM' % (r, w.mean() / 1e6, w.max() / 1e6))
