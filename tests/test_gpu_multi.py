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


def test_one_process_two_devices_deliver_the_full_vectors():
    """``devices=[0, 1]``: one process, one handle per GPU, every shard
    copies into its slice of one pinned host vector; the result is the
    single-device result bit for bit, instance constraints included."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from conftest import assert_values_close
    from opty_b200 import ConstraintCollocator
    from oracle.opty_oracle import OracleCollocator
    for make in (lambda: workloads.n_link_pendulum_torques(4, 203),
                 lambda: workloads.n_link_pendulum(10, 40, seed=7)):
        w = make()
        one = ConstraintCollocator(*w.collocator_args(),
                                   **w.collocator_kwargs(), device=0)
        two = ConstraintCollocator(*w.collocator_args(),
                                   **w.collocator_kwargs(), devices=[0, 1])
        free = w.free(one.num_free)
        con1 = one.generate_constraint_function()(free)
        jac1 = np.array(one.generate_jacobian_function()(free))
        con_f = two.generate_constraint_function()
        jac_f = two.generate_jacobian_function()
        for point in (free, free * 1.01, free):
            con2 = con_f(point)
            jac2 = np.array(jac_f(point))
        assert np.array_equal(con2, con1)
        assert np.array_equal(jac2, jac1)
        r1, c1 = one.jacobian_indices()
        r2, c2 = two.jacobian_indices()
        assert np.array_equal(r1, r2) and np.array_equal(c1, c2)
        orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
        P = orc.P
        nnz = (orc.N - 1) * orc.M * P
        assert_values_close(jac2[:nnz], orc.jacobian(free)[:nnz], row_len=P)
        one.close()
        two.close()
