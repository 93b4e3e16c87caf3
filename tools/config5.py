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
tests/golden/make_cfg5_node0_rows.py), and the Jacobian against directional finite
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


def prepared():
    """True if the module dump and every cubin it names are in the cache."""
    if not os.path.exists(DUMP):
        return False
    try:
        with open(DUMP) as f:
            dump = json.load(f)
        from opty_b200 import build, codegen
        if dump['meta'].get('emitter_version') != codegen.EMITTER_VERSION or \
                dump.get('skeleton_digest') != build._header_digest():
            return False
        paths = [dump['cubin']] + [em['cubin_path'] for em in
                                   dump['meta'].get('extra_modules', ())]
        return all(os.path.getsize(os.path.join(ROOT, p)) > 0 for p in paths)
    except (OSError, ValueError, KeyError):
        return False


def prepare():
    """OPTY_VARIANTS='[["tag", {opts}], ...]' prepares several kernel
    variants after deriving the equations of motion once."""
    import pickle
    import workloads
    global OPTS, TAG, DUMP
    t0 = time.time()
    cache = os.path.join(ROOT, 'opty_b200', '_cache', 'eom_{}.pkl'.format(LINKS))
    if os.path.exists(cache) and LINKS in (20, 50):
        # derived equations of motion kept between runs in the build
        # container (tests/golden/make_sampled_jacobian.py)
        sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
        sys.setrecursionlimit(100000)
        from make_sampled_jacobian import load_workload
        w = load_workload(LINKS, N_FULL, 0)
    else:
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
    pm.meta.pop('entry_kind', None)
    dump = {
        'meta': pm.meta, 'opts': opts, 'num_nodes_full': N_FULL,
        'skeleton_digest': __import__('opty_b200.build', fromlist=['x'])
        ._header_digest(),
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


def full_free_vector(dump):
    """The seeded free vector of the full problem (workloads.n_link_pendulum
    draws the constants first, then the free vector, from one generator)."""
    rng = np.random.default_rng(0)
    for _ in range(dump['rng_param_draws']):
        rng.random()
    return rng.standard_normal(dump['num_free_full'])


def make_handle(dump, N, node_range=None, device=0):
    from opty_b200 import runtime
    cfg = runtime.ColloCfg()
    cfg.abi_version = runtime.ABI_VERSION
    cfg.device = device
    cfg.N = N
    cfg.node_lo, cfg.node_hi = node_range or (0, N - 1)
    for key in ('n', 'q', 'k', 'r', 's', 'pk', 'M', 'P'):
        setattr(cfg, key, dump[key])
    cfg.method = 1
    cfg.out_ring = int(dump['opts']['out_ring'])
    cfg.prefetch_jac = 0
    cfg.con_tail = cfg.jac_tail = 0
    cfg.h = dump['h']
    with open(os.path.join(ROOT, dump['cubin']), 'rb') as f:
        cubin = f.read()
    h = runtime.ColloHandle(cfg, cubin)
    for em in dump['meta'].get('extra_modules', ()):
        with open(os.path.join(ROOT, em['cubin_path']), 'rb') as f:
            h.add_module(f.read())
    h.set_known(None, np.array(dump['params']))
    return h


def run():
    from opty_b200 import runtime
    with open(DUMP) as f:
        dump = json.load(f)
    out = {'tag': TAG, 'opts': OPTS, 'prepare': dump['prepare']}
    n, q, M, P = dump['n'], dump['q'], dump['M'], dump['P']
    free = full_free_vector(dump)
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
        assert rel.max() < 1e-10, rel
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
    groups_meta = dump['meta']['groups'] + [
        g for em in dump['meta'].get('extra_modules', ())
        for g in em['groups']]
    out['group_ops'] = [g['ops'] for g in groups_meta]
    out['peak_live'] = max(g['peak_live'] for g in groups_meta)
    h.close()
    print(json.dumps(out))


def run_sharded():
    """torchrun entry: the 50 000-node problem sharded by nodes over the
    GPUs of the box (BASELINE configs[4]: 8 x B200 with an NCCL all-gather of
    the per-shard residual and Jacobian blocks)."""
    import torch
    import torch.distributed as dist
    from opty_b200.sharding import _CudaArray, gather_vectors, node_shard
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    with open(DUMP) as f:
        dump = json.load(f)
    n, q, M, P = dump['n'], dump['q'], dump['M'], dump['P']
    free = full_free_vector(dump)
    lo, hi = node_shard(N_FULL, rank, world)
    h = make_handle(dump, N_FULL, (lo, hi), local)
    h.upload_free(free)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    h.time_device_evals(2)
    barrier()
    ms = max_over_ranks(min(h.time_device_evals(5) / 5 for _ in range(3)))
    barrier()
    # NCCL all-gather of the shards' blocks straight from the device buffers
    h.eval_device(sync=True)
    bufs = h.device_buffers()
    dev = torch.device('cuda', local)
    con = torch.as_tensor(_CudaArray(bufs['con'], h.con_len, h), device=dev)
    jac = torch.as_tensor(_CudaArray(bufs['jac'], h.jac_len, h), device=dev)
    full_con, full_jac = gather_vectors(con, jac, N_FULL, M, dist)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        full_con, full_jac = gather_vectors(con, jac, N_FULL, M, dist)
    barrier()
    gather_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / 3)
    # every rank's slice of the gathered vector is its own block
    K = M * P
    ok = bool(torch.equal(full_jac[lo * K:hi * K], jac))
    checksum = float(full_jac.sum().item())
    if rank == 0:
        nnf = N_FULL - 1
        B = 8 * ((n + q) * N_FULL + M * nnf + nnf * K)
        print(json.dumps({
            'workload': '{}-link chain, {} midpoint nodes'.format(
                LINKS, N_FULL),
            'tag': TAG, 'n_gpus': world, 'scaling': 'strong',
            'device_ms_per_eval': ms, 'algorithmic_GB': B / 1e9,
            'achieved_GBps_aggregate': B / ms / 1e6,
            'nccl_allgather_ms': gather_ms,
            'allgather_GB_per_rank_out': full_jac.numel() * 8 / 1e9,
            'allgather_busbw_GBps': (full_jac.numel() * 8 / 1e9) *
            (world - 1) / world / (gather_ms * 1e-3),
            'own_block_intact': ok, 'jac_checksum': checksum}), flush=True)
    h.close()
    dist.barrier()
    dist.destroy_process_group()


def profile():
    """A few device-resident evaluations at OPTY_PROFILE_NODES nodes (for
    ncu captures)."""
    with open(DUMP) as f:
        dump = json.load(f)
    N = int(os.environ.get('OPTY_PROFILE_NODES', 10000))
    h = make_handle(dump, N)
    free = np.random.default_rng(0).standard_normal(
        (dump['n'] + dump['q']) * N)
    h.upload_free(free)
    print('ms per eval', h.time_device_evals(3) / 3)
    h.close()


if __name__ == '__main__':
    {'prepare': prepare, 'run': run, 'run_sharded': run_sharded,
     'profile': profile}[sys.argv[1]]()
