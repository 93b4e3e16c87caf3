#!/usr/bin/env python
"""Benchmark of the collocation constraint + Jacobian hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): constraint+Jacobian evals/sec on the 10-link pendulum
at 10 000 midpoint nodes (``configs[1]``); one *eval* = one
``Problem.constraints(free)`` + one ``Problem.jacobian(free)`` at the same
``free``.  One *step* = one eval.

Own arm
    ``value``   K device-resident evals (free vector already in HBM, results
                left in HBM), timed with CUDA events on the launching stream,
                max over ranks.
    ``e2e``     K evals through the public API with HOST arrays: H2D of the
                free vector, kernels, D2H of residuals and Jacobian inside the
                timed region.
    ``roofline``  algorithmic bytes of one fused launch / average launch
                duration, against the measured HBM peak.
    ``cpu_baseline``  the CPU oracle (a port of the reference's generated
                C + loop, bit-identical to it on this workload) on the host
                cores, on a bounded number of evals.

Reference arm (``--impl reference``): the same workload evaluated by the CPU
oracle with all host threads (OpenMP), rank 0 only.

With N > 1 the constraint nodes are sharded: the problem has ``N * 9999``
constraint nodes and rank g evaluates nodes ``[g*9999, (g+1)*9999)`` -- weak
scaling, no data-path collective.  ``value`` is reported in
10k-node-problem evals per second summed over the ranks.
"""

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'constraint+Jacobian evals/sec, 10-link pendulum @ 10k nodes'
UNIT = 'evals/s'
LINKS = 10
NODES_PER_RANK = 10000          # collocation nodes of BASELINE configs[1]
WORKLOAD = ('configs[1]: 10-link inverted pendulum on cart, 10 000 midpoint '
            'nodes, n=M=22, q=1, all constants known, seeded N(0,1) free '
            'vector (SURVEY.md §8d)')
OUT_RING = 4                    # 4 x 83 MB output sets > 126 MB L2


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except (OSError, ValueError, KeyError):
        return 6650.0, 'fallback'


class ClockSampler(object):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed
    regions run (B200_PROFILING.md)."""

    QUERY = ('clocks.sm,clocks.max.sm,power.draw,'
             'clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(device), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax = [], []
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return None
        return {'sm_mhz': float(np.median(sm)),
                'sm_max_mhz': float(np.max(smax)),
                'samples': len(sm), 'reasons': sorted(reasons)}


def build_workload(world):
    import workloads
    n_total = (NODES_PER_RANK - 1) * world + 1
    return workloads.n_link_pendulum(LINKS, n_total)


def algorithmic_bytes(n, q, k, r, s, M, P, N_cols, nodes):
    """Bytes of one fused constraint+Jacobian launch: the
    trajectory columns and free scalars read once, every residual and every
    Jacobian entry (structural zeros included) written once."""
    return 8 * ((n + q + k) * N_cols + r + s + M * nodes + nodes * M * P)


# ---------------------------------------------------------------------------
# CPU arms.  The reference's own implementation (baseline/_ref, installed
# unmodified with pip --target; see baseline/reference.py) is what is timed:
# ``ConstraintCollocator(backend='cython', parallel=True)`` with the
# protocol of BASELINE.md §3.  Only if it cannot be imported does the oracle
# port (oracle/opty_oracle.py, bit-identical results) take its place.
# ---------------------------------------------------------------------------
def _cpu_callables(parallel, threads):
    """Returns ``(con, jac, num_free, workload, kind, effective_threads)``
    for BASELINE configs[1] on the host cores."""
    from baseline import reference as ref
    threads = ref.prepare_environment(threads)
    w = build_workload(1)
    if ref.available():
        try:
            col = ref.collocator(w, parallel)
            con = col.generate_constraint_function()
            jac = col.generate_jacobian_function()
            eff = ref.omp_max_threads() if parallel else 1
            return con, jac, col.num_free, w, 'reference', eff or threads
        except Exception as err:  # pragma: no cover - depends on the host
            sys.stderr.write('reference import/build failed ({}: {}); using '
                             'the oracle port\n'.format(
                                 type(err).__name__, err))
    from oracle.opty_oracle import OracleCollocator
    orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                           parallel=parallel)
    eff = ref.omp_max_threads() if parallel else 1
    return (orc.constraints, lambda f: orc.jacobian(f, copy=False),
            orc.num_free, w, 'port', eff or threads)


def _time_cpu_evals(con, jac, frees, steps, warmup):
    for i in range(max(warmup, 1)):
        con(frees[i % 2])
        jac(frees[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        con(frees[i % 2])
        jac(frees[i % 2])
    return time.perf_counter() - t0


def run_reference_arm(args, rank, world):
    """The reference's CPU path with every host thread it can use; rank 0
    only (the other ranks exit without work)."""
    if rank != 0:
        return
    from baseline import reference as ref
    threads = ref.host_threads()
    con, jac, num_free, w, kind, eff = _cpu_callables(True, threads)
    free = w.free(num_free)
    frees = [free, free + 1e-3]
    dt = _time_cpu_evals(con, jac, frees, args.steps, args.warmup)
    value = args.steps / dt
    what = ("the unmodified reference (baseline/_ref): ConstraintCollocator("
            "backend='cython', parallel=True), generated C + Cython prange "
            "loop" if kind == 'reference' else
            'oracle port of the generated C + node loop (gcc -O2 -fopenmp)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD,
                   'note': 'CPU only: {}; one full 10k-node eval '
                           '(constraints + jacobian at a new point) per '
                           'step'.format(what)},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': eff,
                         'kind': kind,
                         'omp_num_threads': os.environ.get('OMP_NUM_THREADS'),
                         'host_threads': threads,
                         'sample': '{} full evals (constraints + jacobian) '
                                   'of the 10k-node workload'.format(
                                       args.steps)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(budget_s=16.0):
    """The reference (or, failing that, the oracle port) timed on the host
    cores of the GPU box in the same run: OpenMP on all threads and serial.
    Runs in a child process so that OMP_NUM_THREADS is set before libgomp is
    loaded, whatever the parent process has already imported."""
    code = ('import json, sys; sys.path.insert(0, {root!r}); import bench; '
            'print(json.dumps(bench._cpu_baseline_child({budget!r})))'
            .format(root=ROOT, budget=budget_s))
    env = dict(os.environ)
    env.pop('OMP_NUM_THREADS', None)
    proc = subprocess.run([sys.executable, '-c', code], capture_output=True,
                          text=True, env=env, cwd=ROOT)
    for ln in reversed(proc.stdout.strip().splitlines()):
        try:
            return json.loads(ln)
        except ValueError:
            continue
    return {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'failed',
            'sample': (proc.stderr or 'no output')[-400:]}


def _cpu_baseline_child(budget_s):
    from baseline import reference as ref
    threads = ref.host_threads()
    out = {}
    kind = eff = None
    for label, par in (('parallel', True), ('serial', False)):
        con, jac, num_free, w, kind_, eff_ = _cpu_callables(par, threads)
        if par:
            kind, eff = kind_, eff_
        free = w.free(num_free)
        frees = [free, free + 1e-3]
        _time_cpu_evals(con, jac, frees, 0, 2)
        times = []
        t_start = time.perf_counter()
        while True:
            f = frees[len(times) % 2]
            t0 = time.perf_counter()
            con(f)
            t1 = time.perf_counter()
            jac(f)
            t2 = time.perf_counter()
            times.append((t1 - t0, t2 - t1))
            if time.perf_counter() - t_start > budget_s / 2 or \
                    len(times) >= 400:
                break
        tc = np.array([t[0] for t in times])
        tj = np.array([t[1] for t in times])
        out[label] = {
            'evals': len(times),
            'evals_per_s': len(times) / float(tc.sum() + tj.sum()),
            'constraints_ms_best': 1e3 * float(tc.min()),
            'constraints_ms_median': 1e3 * float(np.median(tc)),
            'jacobian_ms_best': 1e3 * float(tj.min()),
            'jacobian_ms_median': 1e3 * float(np.median(tj))}
    par = out['parallel']
    return {'value': par['evals_per_s'], 'unit': UNIT, 'cores': eff,
            'kind': kind, 'host_threads': threads,
            'sample': '{} full evals of the 10k-node workload, '
                      "backend='cython', parallel=True on {} OpenMP threads "
                      '(serial, 1 core: {:.2f} evals/s over {} evals)'.format(
                          par['evals'], eff, out['serial']['evals_per_s'],
                          out['serial']['evals']),
            'serial_value': out['serial']['evals_per_s'],
            'detail': out}


def config5_strong_scaling(rank, world, local_rank, dist, torch):
    """BASELINE configs[4]: the 50-link chain at 50 000 midpoint nodes (8.49 GB
    of residuals + Jacobian per evaluation), its nodes sharded over the GPUs
    of the run (STRONG scaling: the problem does not grow), device resident,
    plus -- for N > 1 -- the NCCL all-gather of the shards' blocks into the
    full vectors on every GPU.  Reported as an extra key next to the headline;
    needs the module prepared by ``__graft_entry__.build()``."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import config5
    from opty_b200.sharding import _CudaArray, gather_vectors, node_shard
    if not config5.prepared():
        return {'skipped': 'config 5 module not prepared'}
    with open(config5.DUMP) as f:
        dump = json.load(f)
    n, q, M, P = dump['n'], dump['q'], dump['M'], dump['P']
    N_full = dump['num_nodes_full']
    free = config5.full_free_vector(dump)
    lo, hi = node_shard(N_full, rank, world)
    h = config5.make_handle(dump, N_full, (lo, hi), local_rank)
    h.upload_free(free)

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=torch.device('cuda', local_rank),
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    h.time_device_evals(2)
    sync()
    ms = max_over_ranks(min(h.time_device_evals(5) / 5 for _ in range(3)))
    nnf = N_full - 1
    K = M * P
    B = 8 * ((n + q) * N_full + M * nnf + nnf * K)
    out = {'workload': 'configs[4]: 50-link chain, {} midpoint nodes, node '
                       'shards over {} GPU(s), strong scaling'.format(
                           N_full, world),
           'device_ms_per_eval': ms, 'algorithmic_GB': B / 1e9,
           'achieved_GBps_aggregate': B / ms / 1e6,
           'evals_per_s': 1e3 / ms}
    if dist is not None:
        h.eval_device(sync=True)
        bufs = h.device_buffers()
        dev = torch.device('cuda', local_rank)
        con = torch.as_tensor(_CudaArray(bufs['con'], h.con_len, h),
                              device=dev)
        jac = torch.as_tensor(_CudaArray(bufs['jac'], h.jac_len, h),
                              device=dev)
        full_con, full_jac = gather_vectors(con, jac, N_full, M, dist)
        sync()
        t0 = time.perf_counter()
        for _ in range(3):
            full_con, full_jac = gather_vectors(con, jac, N_full, M, dist)
        sync()
        gather_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / 3)
        ok = bool(torch.equal(full_jac[lo * K:hi * K], jac))
        out.update({
            'nccl_allgather_ms': gather_ms,
            'allgather_path': 'equal shards: blocks land in the final vector'
            if nnf % world == 0 else
            'ragged shards ({} nodes over {} ranks): padded gather + '
            'concatenation'.format(nnf, world),
            'allgather_busbw_GBps': (full_jac.numel() * 8 / 1e9) *
            (world - 1) / world / (gather_ms * 1e-3),
            'own_block_intact_after_gather': ok})
        del full_con, full_jac
    h.close()
    return out


def run_own_arm(args, rank, world, local_rank):
    import torch
    from opty_b200 import ConstraintCollocator

    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback).')
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(
            'cuda', local_rank))

    numa_cpus = None
    if world > 1:
        # pinned host buffers on the GPU's own socket (one process per GPU)
        from opty_b200.sharding import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local_rank)
    w = build_workload(world)
    nn = NODES_PER_RANK - 1
    node_range = (rank * nn, (rank + 1) * nn)
    col = ConstraintCollocator(
        *w.collocator_args(), **w.collocator_kwargs(), device=local_rank,
        node_range=node_range, cuda_options={'out_ring': OUT_RING})
    con_f = col.generate_constraint_function()
    jac_f = col.generate_jacobian_function()
    ev = col._evaluator
    h = ev.handle
    free = w.free(col.num_free)
    frees = [free, free + 1e-3]
    prog = ev.program

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- device-resident evals ------------------------------------------
    h.upload_free(free)
    h.time_device_evals(max(args.warmup, 3))
    launches0 = h.launch_count()
    barrier()
    ms = h.time_device_evals(args.steps)
    barrier()
    launches = h.launch_count() - launches0
    per_launch_ms = []
    for _ in range(min(args.steps, 50)):
        h.eval_device(sync=True)
        per_launch_ms.append(h.last_kernel_ms())
    if dist is not None:
        t = torch.tensor([ms], device=torch.device('cuda', local_rank),
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps / (ms * 1e-3)

    # ---- end to end through the public API ----------------------------------
    # N = 1: this process' collocator.  N > 1: ONE process (rank 0) drives all
    # N GPUs through ``devices=`` and receives the whole problem's residual
    # and Jacobian vectors in one pinned host buffer each -- what a host-side
    # IPOPT consumes; the other ranks wait.
    h2d_rows = col.num_states + col.num_unknown_input_trajectories
    ranges = getattr(ev, 'd2h_ranges', [(0, prog.K)])
    if world == 1:
        for i in range(max(args.warmup, 3)):
            con_f(frees[i % 2])
            jac_f(frees[i % 2])
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            con = con_f(frees[i % 2])
            jac = jac_f(frees[i % 2])
        barrier()
        e2e_s = time.perf_counter() - t0
        assert con.shape == (prog.M * nn,) and jac.shape == (nn * prog.K,)
        e2e_how = 'Problem-level callables of this process, host arrays'
    else:
        barrier()
        e2e_s = 0.0
        if rank == 0:
            col_all = ConstraintCollocator(
                *w.collocator_args(), **w.collocator_kwargs(),
                devices=list(range(world)), cuda_options={'out_ring': 2})
            con_a = col_all.generate_constraint_function()
            jac_a = col_all.generate_jacobian_function()
            for i in range(max(args.warmup, 3)):
                con_a(frees[i % 2])
                jac_a(frees[i % 2])
            t0 = time.perf_counter()
            for i in range(args.steps):
                con = con_a(frees[i % 2])
                jac = jac_a(frees[i % 2])
            e2e_s = time.perf_counter() - t0
            assert con.shape == (prog.M * nn * world,)
            assert jac.shape == (nn * world * prog.K,)
            col_all.close()
        # (the multi-device collocator made other devices current)
        torch.cuda.set_device(local_rank)
        barrier()
        e2e_how = ('rank 0 drives all {} GPUs (devices=), full residual and '
                   'Jacobian vectors assembled in one pinned host buffer '
                   'each'.format(world))
    if dist is not None:
        t = torch.tensor([e2e_s], device=torch.device('cuda', local_rank),
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * args.steps / e2e_s
    h2d = 8 * (h2d_rows * (nn * world + 1) + col.num_unknown_parameters +
               int(col._variable_duration))
    d2h = 8 * world * (prog.M * nn + nn * sum(e - b for b, e in ranges))

    clocks = sampler.stop() if sampler is not None else None

    extra5 = None
    if not args.no_config5:
        try:
            extra5 = config5_strong_scaling(rank, world, local_rank, dist,
                                            torch)
        except Exception as err:  # the headline must not depend on it
            extra5 = {'failed': '{}: {}'.format(type(err).__name__, err)}

    if rank == 0:
        peak, peak_kind = _peaks()
        bytes_launch = algorithmic_bytes(
            col.num_states, col.num_unknown_input_trajectories,
            col.num_known_input_trajectories, col.num_unknown_parameters,
            int(col._variable_duration), prog.M, prog.P, nn + 1, nn)
        launch_ms = ms / args.steps
        achieved = bytes_launch / (launch_ms * 1e-3) / 1e9
        # DRAM traffic per launch: from an ncu --set full capture of a launch
        # INSIDE the rotating-output loop (profiles/roofline_traffic.json says
        # which); an offline capture, not measured in this run
        traffic = traffic_note = None
        tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    tj = json.load(f)
                traffic = tj.get('dram_bytes_per_launch')
                traffic_note = tj.get('source')
            except (OSError, ValueError):
                traffic = None
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': launch_ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {
                'workload': WORKLOAD,
                'value_is': 'device-resident evaluations (free vector and '
                            'results stay in HBM); the solver-visible rate '
                            'is e2e',
                'nodes_per_gpu': nn, 'parallelism': 'node-shard x{}'.format(
                    world),
                'groups': ev.meta['num_groups'],
                'tile_cols': ev.meta['C'],
                'tile_bufs': ev.meta['tile_bufs'],
                'warps_per_block': ev.meta['warps_per_block'],
                'min_blocks_per_sm': ev.meta['min_blocks_per_sm'],
                'tma_load': ev.meta['tma_load'],
                'tma_store': ev.meta['tma_store'],
                'schedule': ev.meta['schedule'],
                'l2': 'rotating {} device output sets ({:.0f} MB) > 126 MB '
                      'L2'.format(OUT_RING, OUT_RING * 8e-6 * (
                          prog.M * nn + nn * prog.K)),
                'host_affinity': ('GPU-local CPU set ({} cores) per rank'
                                  .format(len(numa_cpus)) if numa_cpus else
                                  'default'),
                'e2e_d2h': 'literal-only Jacobian column ranges are written '
                           'to the pinned buffer once and not re-copied; '
                           'copied columns: {}'.format(ranges),
            },
            'roofline': {
                'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'traffic_source': traffic_note,
                'peak_kind': peak_kind,
                'algorithmic_bytes_per_launch': bytes_launch,
                'kernel': 'opty_colloc_eval',
                'launch_ms_avg': launch_ms,
                'launch_ms_median_isolated': float(np.median(per_launch_ms)),
            },
            'e2e': {'value': e2e_value, 'unit': UNIT,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': 1e3 * e2e_s / args.steps,
                    'how': e2e_how},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'config5_strong_scaling': extra5,
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline()
        print(json.dumps(line), flush=True)
    col.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-config5', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
    else:
        run_own_arm(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
