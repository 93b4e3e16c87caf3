"""BASELINE config 4 (stand-in, see workloads.n_link_pendulum_torques): strong
scaling of a 20 000-node backward-Euler problem with unknown torques, a known
input trajectory, unknown parameters, a free time interval and instance
constraints over the GPUs of one box; one process per GPU, node shards
(opty_b200/sharding.py), NCCL only for the optional device gather.

    python tools/config4_scaling.py prepare                       # build container: fill the module cache
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/config4_scaling.py run              # GPU box

Prints one JSON line per run (rank 0): device-resident evals/s of the whole
problem (max over ranks), per-rank end-to-end evals/s through the Python
callbacks with host buffers, and the time of the NCCL all-gather of the
shards' residual and Jacobian blocks.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import workloads  # noqa: E402

LINKS = int(os.environ.get('OPTY_LINKS', 8))
NODES = int(os.environ.get('OPTY_NODES', 20000))
STEPS = int(os.environ.get('OPTY_STEPS', 300))
OPTS = {'out_ring': 4}


def main():
    mode = sys.argv[1]
    w = workloads.n_link_pendulum_torques(LINKS, NODES)
    if mode == 'prepare':
        from opty_b200 import ConstraintCollocator
        from opty_b200.sharding import node_shard
        for world in (1, 2, 4, 8):
            for rank in (0, world - 1):
                col = ConstraintCollocator(
                    *w.collocator_args(), **w.collocator_kwargs(),
                    node_range=node_shard(NODES, rank, world),
                    cuda_options=OPTS)
                pm = col.prepare_module()
                print(json.dumps({
                    'world': world, 'rank': rank, 'cache_hit': pm.cache_hit,
                    'stats': pm.meta['stats'],
                    'groups': [g['ops'] for g in pm.meta['groups']]}))
        return
    import torch
    import torch.distributed as dist
    from opty_b200.sharding import ShardedCollocator
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    sc = ShardedCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                           rank=rank, world_size=world, device=local,
                           cuda_options=OPTS)
    col = sc.collocator
    h = col._evaluator.handle
    free = w.free(col.num_free)
    frees = [free, free * (1.0 + 1e-6)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    h.upload_free(free)
    h.time_device_evals(10)
    barrier()
    ms = max_over_ranks(h.time_device_evals(STEPS)) / STEPS
    barrier()
    for i in range(5):
        sc.constraints_local(frees[i % 2])
        sc.jacobian_local(frees[i % 2])
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(20, STEPS // 5)
    for i in range(n_e2e):
        sc.constraints_local(frees[i % 2])
        sc.jacobian_local(frees[i % 2])
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / n_e2e)
    gather_ms = None
    if world > 1:
        sc.allgather_device(free)
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            sc.allgather_device()
        barrier()
        gather_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / 10)
    if rank == 0:
        prog = col._evaluator.program
        nn = NODES - 1
        R = (col.num_states + col.num_unknown_input_trajectories +
             col.num_known_input_trajectories)
        B = 8 * (R * NODES + prog.M * nn + nn * prog.K)
        print(json.dumps({
            'workload': 'config-4 stand-in: {}-link pendulum with joint '
                        'torques, {} backward-Euler nodes, n={} q={} k={} r={} '
                        's=1 o={}'.format(
                            LINKS, NODES, col.num_states,
                            col.num_unknown_input_trajectories,
                            col.num_known_input_trajectories,
                            col.num_unknown_parameters,
                            col.num_instance_constraints),
            'n_gpus': world, 'scaling': 'strong', 'M': prog.M, 'P': prog.P,
            'device_ms_per_eval': ms, 'device_evals_per_s': 1e3 / ms,
            'algorithmic_MB': B / 1e6, 'achieved_GBps': B / ms / 1e6,
            'e2e_ms_per_eval_local_shard': e2e_ms,
            'e2e_evals_per_s': 1e3 / e2e_ms,
            'nccl_allgather_eval_plus_gather_ms': gather_ms,
            'groups': col._evaluator.meta['num_groups']}), flush=True)
    sc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
