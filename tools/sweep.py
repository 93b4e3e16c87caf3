"""Development aid: times the device-resident evaluation of BASELINE config 2
for a list of kernel-geometry variants (cuda_options).

    python tools/sweep.py compile   # build container: fill the module cache
    python tools/sweep.py run       # GPU box: time every variant
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import workloads  # noqa: E402
from opty_b200 import ConstraintCollocator  # noqa: E402

VARIANTS = json.load(open(os.path.join(ROOT, 'tools', 'sweep_variants.json')))


def main():
    mode = sys.argv[1]
    w = workloads.n_link_pendulum(10, 10000)
    free = None
    ref = None
    for name, opts in VARIANTS:
        opts = dict(opts)
        for key in ('OPTY_B200_REPL_ROWS', 'OPTY_B200_REPL_TILES',
                    'OPTY_B200_REPL_MODE', 'OPTY_B200_REPL_NODES',
                    'OPTY_B200_NO_PDL'):
            os.environ.pop(key, None)
        for key, val in opts.pop('env', {}).items():
            os.environ[key] = str(val)
        cap = opts.pop('resident_blocks', None)
        if cap:
            os.environ['OPTY_B200_DEBUG_SMEM_FLOOR'] = str(
                (227 * 1024) // cap - 1024)
        else:
            os.environ.pop('OPTY_B200_DEBUG_SMEM_FLOOR', None)
        opts.setdefault('out_ring', 4)
        col = ConstraintCollocator(*w.collocator_args(),
                                   **w.collocator_kwargs(), cuda_options=opts)
        if mode == 'compile':
            t0 = time.time()
            pm = col.prepare_module()
            print(name, 'compiled' if not pm.cache_hit else 'cached',
                  '%.1fs' % (time.time() - t0),
                  [g['ops'] for g in pm.meta['groups']], flush=True)
            continue
        try:
            col.generate_constraint_function()
        except Exception as e:  # noqa: BLE001
            print(name, 'FAILED', str(e)[:200], flush=True)
            continue
        h = col._evaluator.handle
        if free is None:
            free = w.free(col.num_free)
        h.upload_free(free)
        h.time_device_evals(20)
        ms = [h.time_device_evals(200) / 200 for _ in range(3)]
        jac = np.array(col.generate_jacobian_function()(free))
        if ref is None:
            ref = jac
        same = bool(np.array_equal(ref, jac))
        if not same:
            same = 'max|d|/max|ref|=%.2e' % (np.max(np.abs(ref - jac)) /
                                               np.max(np.abs(ref)))
        B = 84551728
        print('%-34s %8.2f us  %7.1f GB/s  bit-equal-to-first=%s' % (
            name, 1e3 * min(ms), B / min(ms) / 1e6, same), flush=True)
        col.close()


if __name__ == '__main__':
    main()
