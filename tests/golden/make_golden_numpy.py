"""Pins the SECOND oracle (``oracle/lambdify_oracle.py``): runs the REFERENCE
itself with ``backend='numpy'`` (``lambdify_matrix``, opty/utils.py:598-636;
plain ``.jacobian``, opty/direct_collocation.py:2757-2758) on small seeded
workloads and stores its outputs.

Build container only:

    python tests/golden/make_golden_numpy.py
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'baseline', 'stubs'))
sys.path.insert(0, '/root/reference')

import workloads  # noqa: E402
from opty.direct_collocation import ConstraintCollocator  # noqa: E402

CASES = {
    'cfg1_pendulum_swing_up_N51_numpy_backend':
        lambda: workloads.pendulum_swing_up(51),
    'cfg3_vyasarayani2011_N101_odd_numpy_backend':
        lambda: workloads.vyasarayani2011(101, seed=5),
    'cfg4_standin_pendulum4_torques_N30_numpy_backend':
        lambda: workloads.n_link_pendulum_torques(4, 30),
    'cfg2_small_pendulum10_N12_numpy_backend':
        lambda: workloads.n_link_pendulum(10, 12, seed=7),
}

if __name__ == '__main__':
    for name, make in CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        w = make()
        col = ConstraintCollocator(*w.collocator_args(),
                                   **w.collocator_kwargs(), backend='numpy')
        free = w.free(col.num_free)
        con = np.array(col.generate_constraint_function()(free))
        jac = np.array(col.generate_jacobian_function()(free))
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, free=free, con=con, jac=jac)
        print(name, con.shape, jac.shape, os.path.getsize(path), flush=True)
