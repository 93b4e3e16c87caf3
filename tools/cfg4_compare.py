"""Development aid: device-resident evaluation time of the BASELINE config-4
stand-in (8 links, 20 000 backward-Euler nodes) with the grid kernel and with
the row-stationary kernel forced onto its odd P.

    python tools/cfg4_compare.py compile | run
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import workloads  # noqa: E402
from opty_b200 import ConstraintCollocator  # noqa: E402

LINKS = int(os.environ.get('OPTY_LINKS', 8))
NODES = int(os.environ.get('OPTY_NODES', 20000))
VARIANTS = [('grid kernel (default for odd P)', {}),
            ('grid kernel, tile-major dispatch', {'tile_major': True}),
            ('grid kernel, 8 groups', {'groups': 8, 'tile_major': False}),
            ('grid kernel, 8 groups, tile-major', {'groups': 8,
                                                   'tile_major': True}),
            ('row-stationary, forced', {'persistent': 'stationary',
                                        'tile_bufs': 1}),
            ('row-stationary, forced, 4 warps',
             {'persistent': 'stationary', 'tile_bufs': 1,
              'warps_per_block': 4})]


def main():
    mode = sys.argv[1]
    w = workloads.n_link_pendulum_torques(LINKS, NODES)
    ref = None
    free = None
    for name, opts in VARIANTS:
        opts = dict(opts, out_ring=4)
        col = ConstraintCollocator(*w.collocator_args(),
                                   **w.collocator_kwargs(), cuda_options=opts)
        if mode == 'compile':
            try:
                pm = col.prepare_module()
            except ValueError as err:
                print(name, 'does not apply:', err)
                continue
            print(name, pm.meta['persistent'], pm.meta['num_groups'],
                  pm.meta.get('smem_bytes'), flush=True)
            continue
        try:
            jac_f = col.generate_jacobian_function()
        except ValueError as err:
            print(name, 'does not apply:', str(err)[:80])
            continue
        h = col._evaluator.handle
        if free is None:
            free = w.free(col.num_free)
        jac = np.array(jac_f(free))
        h.time_device_evals(20)
        ms = min(h.time_device_evals(100) / 100 for _ in range(3))
        if ref is None:
            ref = jac
        prog = col._evaluator.program
        nn = NODES - 1
        B = 8 * (nn * prog.M * prog.P + prog.M * nn + prog.R * NODES)
        print(json.dumps({'variant': name, 'us_per_eval': 1e3 * ms,
                          'GBps': B / ms / 1e6,
                          'max_rel_diff_to_first': float(
                              np.max(np.abs(jac - ref)) /
                              np.max(np.abs(ref)))}), flush=True)
        col.close()


if __name__ == '__main__':
    main()
