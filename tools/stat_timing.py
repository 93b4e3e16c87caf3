"""Development aid: summarises the per-block timing lines the row-stationary
kernel prints with cuda option debug_nostore & 2."""
import collections
import re
import sys


def main(path, nslots=148):
    lines = [l for l in open(path) if l.startswith('T slot')][-nslots:]
    rec = []
    for l in lines:
        m = re.match(r'T slot (\d+) t0 (\d+) total (\d+) items (\d+) : (.*)', l)
        items = [tuple(int(x) for x in it.split()) for it in m[5].split('|')]
        rec.append((int(m[1]), int(m[2]), int(m[3]), int(m[4]), items))
    tmin = min(r[1] for r in rec)
    print(path, 'start spread', max(r[1] for r in rec) - tmin, 'ns; last end',
          max(r[1] + r[2] for r in rec) - tmin, 'block total avg %.0f' %
          (sum(r[2] for r in rec) / len(rec)), 'max', max(r[2] for r in rec),
          'min', min(r[2] for r in rec))
    d, first = collections.defaultdict(list), collections.defaultdict(list)
    for r in rec:
        for k, (g, dt) in enumerate(r[4]):
            if g >= 0:
                (first if k == 0 else d)[g].append(dt)
    for g in sorted(set(d) | set(first)):
        a, f = d.get(g, []), first.get(g, [])
        print('  group', g, 'later items n=%d avg %.0f min %d max %d' % (
            len(a), sum(a) / max(len(a), 1), min(a or [0]), max(a or [0])),
            '| first items n=%d avg %.0f' % (len(f), sum(f) / max(len(f), 1)))
    rec.sort(key=lambda r: -r[2])
    for r in rec[:4] + rec[-3:]:
        print('   slot', r[0], 'total', r[2], 'items', r[3], [x for x in r[4] if x[0] >= 0])


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
