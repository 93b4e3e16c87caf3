"""BASELINE config 5 on one GPU: 50-link chain, 50 000 midpoint nodes
(n = M = 102, P = 206, 21 012 Jacobian entries per node, 8.4 GB of Jacobian
values per evaluation).

    python tools/config5.py prepare   # build container: derive (3 min), lower, emit, nvcc (~10 min);
                                      # fills the module cache and writes a SymPy-free problem dump
    python tools/config5.py run       # GPU box: correctness checks at N = 2 000, timing at N = 50 000,
                                      # straight from the dump (no SymPy work on the GPU box)

OPTY_OPTS='{"warps_per_block": 8, ...}' selects kernel options; OPTY_TAG names the dump.

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

OPTS = {'prefetch_jacobian': False, 'd2h_skip_constants': False}
OPTS.update(json.loads(os.environ.get('OPTY_OPTS', '{}')))
TAG = os.environ.get('OPTY_TAG', 'default')
LINKS = int(os.environ.get('OPTY_LINKS', 50))
N_FULL = int(os.environ.get('OPTY_NODES', 50000))
N_CHECK = 2000
DUMP = os.path.join(ROOT, 'opty_b200', '_cache',
                    'config5_{}_{}.json'.format(LINKS, TAG))


def prepare():
    """OPTY_VARIANTS='[["tag", {opts}], ...]' prepares several kernel
    variants after deriving the equations of motion once."""
    import workloads
    global OPTS, TAG, DUMP
    t0 = time.time()
    w = workloads.n_link_pendulum(LINKS, N_FULL)
    derive_s = time.time() - t0
    variants = json.loads(os.environ.get('OPTY_VARIANTS', 'null'))
    if not variants:
        variants = [[TAG, {}]]
    base = dict(OPTS)
    for tag, extra in variants:
        OPTS = dict(base)
        OPTS.update(extra)
        TAG = tag
        DUMP = os.path.join(ROOT, 'opty_b200', '_cache',
                            'config5_{}_{}.json'.format(LINKS, TAG))
        prepare_one(w, derive_s)


def prepare_one(w, derive_s):
    from opty_b200 import ConstraintCollocator
    from opty_b200.direct_collocation import DEFAULT_CUDA_OPTIONS
    out = {'tag': TAG, 'derive_s': derive_s}
    t0 = time.time()
    col = ConstraintCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                               cuda_options=OPTS)
    out['collocator_s'] = time.time() - t0
    t0 = time.time()
    pm = col.prepare_module()
    out['prepare_s'] = time.time() - t0
    out['groups'] = len(pm.parts)
    out['derived_rows'] = len(pm.derived)
    out['cache_hit'] = pm.cache_hit
    out['stats'] = pm.meta['stats']
    opts = dict(DEFAULT_CUDA_OPTIONS)
    opts.update(OPTS)
    for em in pm.meta.get('extra_modules', ()):
        em['cubin_path'] = os.path.relpath(em['cubin_path'], ROOT)
    dump = {
        'meta': pm.meta, 'opts': opts,
        'cubin': os.path.relpath(pm.cubin_path, ROOT),
        'n': col.num_states, 'q': col.num_unknown_input_trajectories,
        'k': col.num_known_input_trajectories,
        'r': col.num_unknown_parameters, 's': int(col._variable_duration),
        'pk': col.num_known_parameters, 'M': pm.program.M, 'P': pm.program.P,
        'h': float(col.node_time_interval),
        'params': [float(col.known_parameter_map[p])
                   for p in col.known_parameters],
        'num_free_full': col.num_free,
        'rng_param_draws': sum(1 for p in w.known_parameter_map
                               if p.name != 'g'),
        'prepare': out,
    }
    with open(DUMP, 'w') as f:
        json.dump(dump, f)
    print(json.dumps(out), flush=True)


def make_handle(dump, N):
    from opty_b200 import runtime
    from opty_b200.direct_collocation import fill_kernel_config
    cfg = runtime.ColloCfg()
    fill_kernel_config(cfg, dump['meta'], dump['opts'])
    cfg.device = 0
    cfg.N = N
    cfg.node_lo, cfg.node_hi = 0, N - 1
    for key in ('n', 'q', 'k', 'r', 's', 'pk', 'M', 'P'):
        setattr(cfg, key, dump[key])
    cfg.method = 1
    cfg.con_tail = cfg.jac_tail = 0
    cfg.h = dump['h']
    with open(os.path.join(ROOT, dump['cubin']), 'rb') as f:
        cubin = f.read()
    h = runtime.ColloHandle(cfg, cubin)
    for em in dump['meta'].get('extra_modules', ()):
        with open(os.path.join(ROOT, em['cubin_path']), 'rb') as f:
            s0, s1 = em['segment_range']
            h.add_module(f.read(), s0, s1 - s0, em['num_groups'])
    if dump['meta']['const_runs']:
        h.set_const_runs(dump['meta']['const_runs'], dump['meta']['const_lit'],
                         dump['meta']['const_inv'])
    h.set_known(None, np.array(dump['params']))
    return h


def run():
    from opty_b200 import runtime
    with open(DUMP) as f:
        dump = json.load(f)
    out = {'tag': TAG, 'opts': OPTS, 'prepare': dump['prepare']}
    n, q, M, P = dump['n'], dump['q'], dump['M'], dump['P']
    rng = np.random.default_rng(0)
    for _ in range(dump['rng_param_draws']):
        rng.random()
    free = rng.standard_normal(dump['num_free_full'])
    n_rows = n + q
    # ---- correctness at N_CHECK (the generated module does not depend on N)
    fc = np.concatenate([free[j * N_FULL:j * N_FULL + N_CHECK]
                         for j in range(n_rows)])
    t0 = time.time()
    hc = make_handle(dump, N_CHECK)
    out['handle_setup_s'] = time.time() - t0
    nn = N_CHECK - 1
    con = hc.constraints(fc).copy()
    jac = np.array(hc.jacobian(fc))
    gold_path = os.path.join(ROOT, 'tests', 'golden',
                             'cfg5_pendulum50_node0_rows.npz')
    if LINKS == 50 and os.path.exists(gold_path):
        gold = np.load(gold_path)
        assert np.array_equal(gold['free_head'], free[:8])
        got = con.reshape(M, nn)[gold['rows'], 0]
        rel = np.abs(got - gold['values']) / np.abs(gold['values'])
        out['residual_rows_checked'] = gold['rows'].tolist()
        out['residual_max_rel_err_vs_sympy_evalf'] = float(rel.max())
        assert rel.max() < 1e-9, rel
    rows, cols = runtime.jacobian_indices(0, N_CHECK, 0, nn, n, q, 0, 0, M, 1)
    d = np.random.default_rng(2).standard_normal(fc.size)
    eps = 1e-6
    fd = (hc.constraints(fc + eps * d).copy() -
          hc.constraints(fc - eps * d).copy()) / (2 * eps)
    jv = np.bincount(rows, weights=jac * d[cols], minlength=len(con))
    out['fd_check_max_abs_over_max'] = float(np.max(np.abs(fd - jv)) /
                                             np.max(np.abs(jv)))
    assert out['fd_check_max_abs_over_max'] < 1e-5
    hc.close()

    # ---- timing at N_FULL, device resident
    h = make_handle(dump, N_FULL)
    h.upload_free(free)
    h.time_device_evals(2)
    ms = [h.time_device_evals(5) / 5 for _ in range(3)]
    nnf = N_FULL - 1
    bytes_launch = 8 * (n_rows * N_FULL + M * nnf + nnf * M * P)
    out['ms_per_eval'] = min(ms)
    out['algorithmic_GB'] = bytes_launch / 1e9
    out['achieved_GBps'] = bytes_launch / (min(ms) * 1e-3) / 1e9
    out['groups'] = dump['meta']['num_groups']
    out['group_ops'] = [g['ops'] for g in dump['meta']['groups']] + [
        g['ops'] for em in dump['meta'].get('extra_modules', ())
        for g in em['groups']]
    h.close()
    print(json.dumps(out))


if __name__ == '__main__':
    {'prepare': prepare, 'run': run}[sys.argv[1]]()
