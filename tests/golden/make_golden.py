"""Generates the golden fixtures in this directory by running the REFERENCE
(csu-hmc/opty, /root/reference, backend='cython') on the seeded workloads of
``workloads.py``.

Run in the build container only (the reference is not available on the GPU
box):

    CC=/usr/bin/gcc LDSHARED="/usr/bin/gcc -shared" \
        python tests/golden/make_golden.py

The reference imports ``cyipopt`` at module level (opty/direct_collocation.py:
10); ``oracle/_stubs/cyipopt`` stands in for it -- the constraint / Jacobian
path never calls IPOPT.

Small workloads are stored completely.  For BASELINE config 2 (10-link
pendulum, 10 000 nodes; 81 MB of Jacobian values) the fixture holds the values
of a sample of nodes plus SHA-256 digests of the complete arrays.
"""

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_stubs'))
sys.path.insert(0, '/root/reference')
os.environ.setdefault('CC', '/usr/bin/gcc')
os.environ.setdefault('LDSHARED', '/usr/bin/gcc -shared')

import workloads  # noqa: E402
from opty.direct_collocation import ConstraintCollocator  # noqa: E402


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def reference_outputs(w, **extra):
    col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                               backend='cython', **extra)
    con = col.generate_constraint_function()
    jac = col.generate_jacobian_function()
    free = w.free(col.num_free)
    c = np.array(con(free))
    j = np.array(jac(free))
    rows, cols = col.jacobian_indices()
    return col, free, c, j, rows.astype(np.int64), cols.astype(np.int64)


def save_full(name, w):
    col, free, c, j, rows, cols = reference_outputs(w)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, free=free, con=c, jac=j, rows=rows, cols=cols)
    print(name, 'free', free.shape, 'con', c.shape, 'nnz', j.shape,
          os.path.getsize(path), 'bytes')


def save_sampled(name, w, num_sample=48):
    col, free, c, j, rows, cols = reference_outputs(w)
    nn = col.num_collocation_nodes - 1
    M = col.num_eom
    K = len(j) // nn
    rng = np.random.default_rng(12345)
    nodes = np.unique(np.concatenate((
        [0, 1, 31, 32, 33, nn - 2, nn - 1],
        rng.integers(0, nn, num_sample))))
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(
        path, nodes=nodes,
        con=c.reshape(M, nn)[:, nodes],
        jac=j.reshape(nn, K)[nodes],
        rows=rows.reshape(nn, K)[nodes], cols=cols.reshape(nn, K)[nodes],
        free_sha256=digest(free), con_sha256=digest(c), jac_sha256=digest(j),
        rows_sha256=digest(rows), cols_sha256=digest(cols),
        num_free=col.num_free, nnz=len(j))
    print(name, 'sampled nodes', len(nodes), 'nnz', len(j),
          os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    only = sys.argv[1:]
    # PYTHONHASHSEED does not matter here: every instance constraint of the
    # fixtures has a single function atom (last hard part)
    if not only or 'cfg1_pendulum_swing_up_N51' in only:
        save_full('cfg1_pendulum_swing_up_N51', workloads.pendulum_swing_up(51))
    if not only or 'cfg3_vyasarayani2011_N5000' in only:
        save_full('cfg3_vyasarayani2011_N5000', workloads.vyasarayani2011(5000))
    if not only or 'cfg3_vyasarayani2011_N101_odd' in only:
        save_full('cfg3_vyasarayani2011_N101_odd',
              workloads.vyasarayani2011(101, seed=5))
    if not only or 'cfg4_standin_pendulum4_torques_N200' in only:
        save_full('cfg4_standin_pendulum4_torques_N200',
              workloads.n_link_pendulum_torques(4, 200))
    if not only or 'cfg2_small_pendulum10_N40' in only:
        save_full('cfg2_small_pendulum10_N40',
              workloads.n_link_pendulum(10, 40, seed=7))
    if not only or 'cfg4_periodic_pendulum4_N200' in only:
        # instance constraints with two function atoms: the order of their
        # Jacobian entries follows Python's set iteration
        # (opty/direct_collocation.py:2244, 2264), i.e. the hash seed of the
        # generating process.  Consumers compare the instance part as a set
        # of (row, col, value) triplets.
        save_full('cfg4_periodic_pendulum4_N200',
                  workloads.n_link_pendulum_periodic(4, 200))
    if not only or 'cfg2_pendulum10_N10000' in only:
        save_sampled('cfg2_pendulum10_N10000',
                 workloads.n_link_pendulum(10, 10000))
