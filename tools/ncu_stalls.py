"""Development aid: aggregates the per-instruction stall samples of an
``ncu --page source --csv`` export by SASS opcode.

    ncu -i prof.ncu-rep --page source --csv > prof_source.csv
    python tools/ncu_stalls.py prof_source.csv
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
byop = collections.defaultdict(collections.Counter)
opcount = collections.Counter()
opexec = collections.Counter()
for r in data:
    src = r[ix['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else '?'
    opcount[op] += 1
    opexec[op] += int(r[ix['Instructions Executed']])
    for s in stalls:
        v = int(r[ix[s]] or 0)
        tot[s] += v
        byop[op][s] += v
T = sum(tot.values())
print('total samples', T)
for s, v in tot.most_common():
    print('%-28s %6d %5.1f%%' % (s, v, 100 * v / T))
print()
print('%-10s %6s %9s %7s  top stalls' % ('op', 'static', 'executed',
                                          'samples'))
for op, c in sorted(byop.items(), key=lambda kv: -sum(kv[1].values()))[:24]:
    sv = sum(c.values())
    print('%-10s %6d %9d %7d  %s' % (
        op, opcount[op], opexec[op], sv,
        ', '.join('%s=%d' % (k[6:], v) for k, v in c.most_common(4))))
print('total static', len(data), 'executed', sum(opexec.values()))
