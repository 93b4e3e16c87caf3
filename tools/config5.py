"""BASELINE config 5 on one GPU: 50-link chain, 50 000 midpoint nodes
(n = M = 102, P = 206, 21 012 Jacobian entries per node, 8.4 GB of Jacobian
values per evaluation).

    python tools/config5.py prepare   # build container: derive, emit, nvcc (~9 min), fills the module cache
    python tools/config5.py run       # GPU box: correctness checks at N = 2 000, timing at N = 50 000

Checks (the CPU oracle cannot be built for this model in any reasonable time,
SURVEY.md §8d): residuals of a few equations at node 0 against SymPy
arbitrary-precision ``evalf`` of the discrete EOM (fixture made by
tools/config5_reference_rows.py), and the Jacobian against directional finite
differences of the residuals.
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

OPTS = {'prefetch_jacobian': False, 'd2h_skip_constants': False}
OPTS.update(json.loads(os.environ.get('OPTY_OPTS', '{}')))
N_FULL, N_CHECK = 50000, 2000


def main():
    mode = sys.argv[1]
    out = {}
    t0 = time.time()
    w = workloads.n_link_pendulum(50, N_FULL)
    out['derive_s'] = time.time() - t0
    t0 = time.time()
    col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                               cuda_options=OPTS)
    out['collocator_s'] = time.time() - t0
    if mode == 'prepare':
        t0 = time.time()
        pm = col.prepare_module()
        out['prepare_s'] = time.time() - t0
        out['groups'] = len(pm.parts)
        out['derived_rows'] = len(pm.derived)
        out['cache_hit'] = pm.cache_hit
        out['stats'] = pm.meta['stats']
        print(json.dumps(out))
        return

    free = w.free(col.num_free)
    # ---- correctness at N_CHECK (same generated module: it does not depend on N)
    wc = workloads.n_link_pendulum(50, N_CHECK)
    wc.eom, wc.states, wc.known_parameter_map = w.eom, w.states, w.known_parameter_map
    cc = ConstraintCollocator(*wc.collocator_args(), **wc.collocator_kwargs(),
                              cuda_options=OPTS)
    n_rows = col.num_states + col.num_unknown_input_trajectories
    fc = np.concatenate([free[j * N_FULL:j * N_FULL + N_CHECK]
                         for j in range(n_rows)])
    t0 = time.time()
    con_f = cc.generate_constraint_function()
    jac_f = cc.generate_jacobian_function()
    out['evaluator_setup_s'] = time.time() - t0
    out['module_cache_hit'] = cc._evaluator.cache_hit
    con = con_f(fc)
    jac = np.array(jac_f(fc))
    M = cc.num_eom
    nn = N_CHECK - 1
    gold_path = os.path.join(ROOT, 'tests', 'golden',
                             'cfg5_pendulum50_node0_rows.npz')
    if os.path.exists(gold_path):
        gold = np.load(gold_path)
        assert np.array_equal(gold['free_head'], free[:8])
        got = con.reshape(M, nn)[gold['rows'], 0]
        rel = np.abs(got - gold['values']) / np.abs(gold['values'])
        out['residual_rows_checked'] = gold['rows'].tolist()
        out['residual_max_rel_err_vs_sympy_evalf'] = float(rel.max())
        assert rel.max() < 1e-9, rel
    rows, cols = cc.jacobian_indices()
    rng = np.random.default_rng(2)
    d = rng.standard_normal(fc.size)
    eps = 1e-6
    fd = (con_f(fc + eps * d) - con_f(fc - eps * d)) / (2 * eps)
    jv = np.bincount(rows, weights=jac * d[cols], minlength=len(con))
    out['fd_check_max_abs_over_max'] = float(np.max(np.abs(fd - jv)) /
                                             np.max(np.abs(jv)))
    assert out['fd_check_max_abs_over_max'] < 1e-5
    cc.close()

    # ---- timing at N_FULL, device resident
    col.generate_constraint_function()
    h = col._evaluator.handle
    h.upload_free(free)
    h.time_device_evals(2)
    ms = [h.time_device_evals(5) / 5 for _ in range(3)]
    P = col._evaluator.program.P
    nnf = N_FULL - 1
    bytes_launch = 8 * (n_rows * N_FULL + M * nnf + nnf * M * P)
    out['ms_per_eval'] = min(ms)
    out['algorithmic_GB'] = bytes_launch / 1e9
    out['achieved_GBps'] = bytes_launch / (min(ms) * 1e-3) / 1e9
    out['groups'] = col._evaluator.meta['num_groups']
    col.close()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
