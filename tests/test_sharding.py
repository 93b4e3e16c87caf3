"""Host-side logic of the multi-GPU path (node sharding + gathers), run with
world size 2 on the ``gloo`` backend.  Shard values come from the CPU oracle
(sliced to each rank's node range), so no GPU is needed."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import workloads
from opty_b200 import sharding
from oracle.opty_oracle import OracleCollocator


def test_node_shard_partitions_all_nodes():
    for N, world in ((10000, 1), (10000, 8), (102, 2), (51, 4), (9, 8)):
        shards = sharding.all_shards(N, world)
        assert shards[0][0] == 0 and shards[-1][1] == N - 1
        for (a0, a1), (b0, b1) in zip(shards, shards[1:]):
            assert a1 == b0 and a1 > a0
        sizes = [hi - lo for lo, hi in shards]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.node_shard(5, 0, 8)
    with pytest.raises(ValueError):
        sharding.node_shard(50, 3, 2)


def test_assemble_is_inverse_of_slicing():
    rng = np.random.default_rng(0)
    M, nn, K = 3, 17, 12
    con = rng.standard_normal((M, nn))
    jac = rng.standard_normal((nn, K))
    shards = sharding.all_shards(nn + 1, 4)
    cb = [con[:, lo:hi].ravel() for lo, hi in shards]
    jb = [jac[lo:hi].ravel() for lo, hi in shards]
    assert np.array_equal(sharding.assemble_constraints(cb, shards, M),
                          con.ravel())
    assert np.array_equal(sharding.assemble_jacobian(jb), jac.ravel())


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, tmpdir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        w = workloads.vyasarayani2011(N, seed=5)
        orc = OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())
        free = w.free(orc.num_free)
        con = orc.constraints(free)
        jac = orc.jacobian(free)
        M, nn = orc.M, N - 1
        K = M * orc.P
        lo, hi = sharding.node_shard(N, rank, world)
        # what this rank's GPU would have produced
        local_con = np.ascontiguousarray(con.reshape(M, nn)[:, lo:hi]).ravel()
        local_jac = jac.reshape(nn, K)[lo:hi].ravel()
        full_con, full_jac = sharding.gather_vectors(
            torch.from_numpy(local_con), torch.from_numpy(local_jac.copy()),
            N, M, dist)
        ok = (np.array_equal(full_con.numpy(), con) and
              np.array_equal(full_jac.numpy(), jac))
        np.save(os.path.join(tmpdir, 'ok{}.npy'.format(rank)),
                np.array([ok, lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('N', [101, 102])
def test_gather_vectors_world_size_2_gloo(tmp_path, N):
    # build the oracle's C once in the parent so both ranks hit the cache
    w = workloads.vyasarayani2011(N, seed=5)
    OracleCollocator(*w.collocator_args(), **w.collocator_kwargs())._loops()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, N, str(tmp_path)), nprocs=2, join=True)
    nn = N - 1
    seen = []
    for rank in range(2):
        ok, lo, hi = np.load(str(tmp_path / 'ok{}.npy'.format(rank)))
        assert ok == 1
        seen.append((int(lo), int(hi)))
    assert seen[0][0] == 0 and seen[0][1] == seen[1][0] and seen[1][1] == nn
