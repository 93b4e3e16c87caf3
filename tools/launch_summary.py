"""Development aid: per-kernel summary of an ncu --csv launch list."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            h, start = r, i
            break
    kn, mn, mv = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
    d = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[start + 1:]:
        if len(r) > mv:
            d[r[kn]][r[mn]].append(float(r[mv].replace(',', '')))
    for k, ms in d.items():
        print(k[:40], ' '.join('%s n=%d avg=%.1f min=%.1f' % (
            m.split('.')[0], len(v), sum(v) / len(v), min(v)) for m, v in ms.items()))


if __name__ == '__main__':
    for p in sys.argv[1:]:
        print(p)
        main(p)
