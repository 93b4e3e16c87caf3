"""Golden fixtures for the LARGE models (20- and 50-link chains), where the C
oracle cannot be built in reasonable time (the reference needs
>10 min of SymPy + >7 min of gcc for the 50-link Jacobian alone): sampled
Jacobian entries and residuals evaluated by the second oracle
(``oracle/lambdify_oracle.py`` -- ``sm.diff`` of one discrete EOM row by one
``wrt`` symbol, then lambdify), both in NumPy float64 (what the reference's
``backend='numpy'`` computes) and in 40-digit mpmath arithmetic (the exact
value, correctly rounded).

Build container only (the derivation of the 50-link equations of motion takes
~3 min; a pickle of them is kept in /tmp/eomcache between runs):

    python tests/golden/make_sampled_jacobian.py 20
    python tests/golden/make_sampled_jacobian.py 50
"""

import os
import pickle
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.setrecursionlimit(100000)

import workloads  # noqa: E402
from oracle.lambdify_oracle import LambdifyOracle  # noqa: E402

# links -> (workload nodes, seed, nodes of the checked problem, fixture name)
CASES = {
    20: (2000, 9, 2000, 'pendulum20_N2000_sampled_entries'),
    50: (50000, 0, 2000, 'cfg5_pendulum50_sampled_entries'),
}
PAIRS = 160         # (row, col) pairs, each evaluated at NODES_PER_PAIR nodes
NODES_PER_PAIR = 2


def load_workload(links, N, seed):
    cache = os.path.join(ROOT, 'opty_b200', '_cache', 'eom_{}.pkl'.format(links))
    if not os.path.exists(cache):
        cache = '/tmp/eomcache/eom_{}.pkl'.format(links)
    w = None
    if os.path.exists(cache):
        # the constants and the free vector are re-drawn from the seed; only
        # the symbolic derivation is skipped
        import sympy as sm
        import sympy.physics.mechanics as me
        me.dynamicsymbols._t = sm.Symbol('t')
        with open(cache, 'rb') as f:
            eom, states, pm = pickle.load(f)
        from collections import OrderedDict
        rng = np.random.default_rng(seed)
        par_map = OrderedDict()
        for sym, _ in pm:
            par_map[sym] = 9.81 if sym.name == 'g' else 0.5 + rng.random()
        assert [v for _, v in pm] == list(par_map.values())
        w = workloads.Workload('pendulum{}_N{}'.format(links, N), eom,
                               list(states), N, 0.001, 'midpoint',
                               known_parameter_map=par_map, seed=seed,
                               free=lambda nf: rng.standard_normal(nf))
    else:
        w = workloads.n_link_pendulum(links, N, seed=seed)
    return w


def main(links):
    N_full, seed, N_check, name = CASES[links]
    t0 = time.time()
    w = load_workload(links, N_full, seed)
    print('equations of motion', time.time() - t0, flush=True)
    n = len(w.states)
    free_full = w.free((n + 1) * N_full)
    # the checked problem: the first N_check columns of every trajectory row
    # (the generated module does not depend on N; tools/config5.py does the
    # same)
    free = np.concatenate([free_full[j * N_full:j * N_full + N_check]
                           for j in range(n + 1)])
    w.num_nodes = N_check
    t0 = time.time()
    lam = LambdifyOracle(*w.collocator_args(), **w.collocator_kwargs())
    print('discretised', time.time() - t0, flush=True)
    M, P, nn = lam.M, lam.P, N_check - 1
    rng = np.random.default_rng(links)
    pairs = []
    # the kinematic rows (q' - u) have literal partials; sample mostly the
    # dynamic rows and a few kinematic ones
    dyn_rows = list(range(M // 2, M))
    kin_rows = list(range(0, M // 2))
    while len(pairs) < PAIRS:
        row = int(rng.choice(dyn_rows if len(pairs) >= 4 else kin_rows))
        col = int(rng.integers(P))
        if (row, col) in pairs or not lam.structural_nonzero(row, col):
            continue
        pairs.append((row, col))
    entries = []
    for row, col in pairs:
        for node in rng.choice(nn, NODES_PER_PAIR, replace=False):
            entries.append((int(node), row, col))
    t0 = time.time()
    exact = lam.jacobian_entries(free, entries, dps=40)
    print('jacobian entries, mpmath', time.time() - t0, flush=True)
    t0 = time.time()
    f64 = lam.jacobian_entries(free, entries)
    print('jacobian entries, numpy', time.time() - t0, flush=True)
    res_entries = sorted({(e[0], e[1]) for e in entries})
    t0 = time.time()
    res_exact = lam.residual_entries(free, res_entries, dps=40)
    res_f64 = lam.residual_entries(free, res_entries)
    print('residuals', time.time() - t0, flush=True)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(
        path, links=links, num_nodes=N_check, num_nodes_full=N_full,
        seed=seed, entries=np.array(entries), jac_exact=exact, jac_f64=f64,
        res_entries=np.array(res_entries), res_exact=res_exact,
        res_f64=res_f64, free_head=free[:8], free_sum=float(free.sum()))
    rel = np.abs(f64 - exact) / np.abs(exact)
    print(name, len(entries), 'entries; float64 lambdify vs exact: max rel',
          rel.max(), 'median', np.median(rel), flush=True)


if __name__ == '__main__':
    main(int(sys.argv[1]))
