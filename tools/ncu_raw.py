"""Development aid: prints selected columns of `ncu --page raw --csv` files side by side."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']


def main(paths, extra):
    tabs = []
    for p in paths:
        rows = [r for r in csv.reader(open(p)) if len(r) > 10]
        tabs.append(dict(zip(rows[0], rows[2])))
    keys = KEYS + [k for k in tabs[0] if any(e in k for e in extra)]
    for k in keys:
        print('%-90s' % k[:90], ' '.join('%14s' % t.get(k, '-')[:14] for t in tabs))


if __name__ == '__main__':
    paths = [a for a in sys.argv[1:] if a.endswith('.csv')]
    extra = [a for a in sys.argv[1:] if not a.endswith('.csv')] or ['issue_stalled', 'warp_issue_stalled']
    main(paths, extra)
