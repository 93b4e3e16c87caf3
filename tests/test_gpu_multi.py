"""Multi-GPU path: two ranks, one GPU each, NCCL.  Skipped on boxes with a
single GPU (the world-size-2 host logic is covered on CPU with gloo in
tests/test_sharding.py)."""

import os
import socket

import numpy as np
import pytest
import torch

import workloads

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    from conftest import assert_values_close
    from opty_b200.sharding import ShardedCollocator
    from oracle.opty_oracle import OracleCollocator
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        w = workloads.n_link_pendulum_torques(4, 203)
        sc = ShardedCollocator(*w.collocator_args(), **w.collocator_kwargs(),
                               rank=rank, world_size=world, device=rank)
        col = sc.collocator
        free = w.free(col.num_free)
        orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
        ocon, ojac = orc.constraints(free), orc.jacobian(free)
        M, nn = orc.M, orc.N - 1
        P = orc.P
        # NCCL all-gather straight from the shards' device buffers
        con_d, jac_d = sc.allgather_device(free)
        assert con_d.is_cuda and jac_d.is_cuda
        assert_values_close(con_d.cpu().numpy(), ocon[:M * nn])
        assert_values_close(jac_d.cpu().numpy(), ojac[:nn * M * P], row_len=P)
        # host-staged gather incl. the instance-constraint tails
        con_h, jac_h = sc.gather_to_host(free)
        assert con_h.shape == ocon.shape and jac_h.shape == ojac.shape
        assert_values_close(con_h[:M * nn], ocon[:M * nn])
        assert_values_close(jac_h[:nn * M * P], ojac[:nn * M * P], row_len=P)
        np.testing.assert_allclose(con_h[M * nn:], ocon[M * nn:], rtol=1e-13)
        np.testing.assert_allclose(jac_h[nn * M * P:], ojac[nn * M * P:],
                                   rtol=1e-13)
        # shard structure = slice of the global structure
        rows, cols = sc.jacobian_indices_local()
        orows, ocols = orc.jacobian_indices()
        lo, hi = sc.node_range
        K = M * P
        assert np.array_equal(rows, orows[lo * K:hi * K])
        assert np.array_equal(cols, ocols[lo * K:hi * K])
        sc.close()
        open(os.path.join(tmpdir, 'ok{}'.format(rank)), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_two_gpu_node_shards_with_nccl_allgather(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    from oracle.opty_oracle import OracleCollocator
    w = workloads.n_link_pendulum_torques(4, 203)
    OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())._loops()
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2,
             join=True)
    assert (tmp_path / 'ok0').exists() and (tmp_path / 'ok1').exists()
