"""world_size-2 `gloo` test of the data-parallel host logic (texturemixer_b200.parallel):
contiguous batch sharding and the flat-bucket SUM all-reduce + 1/N scaling semantic of
tfutil.py:326-344, against the oracle's tower-sum restatement.  CPU only."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from oracle import optim_ref as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    from texturemixer_b200 import parallel
    parallel.init_from_env(backend='gloo')
    assert parallel.world_size() == world and parallel.rank() == rank
    b, e = parallel.shard_bounds(8)
    rng = np.random.RandomState(100 + rank)
    g1 = torch.from_numpy(rng.randn(1000).astype(np.float32))
    g2 = torch.from_numpy(rng.randn(64).astype(np.float32))
    empty = torch.zeros(0)
    n = parallel.allreduce_sum_([g1, g2, empty])
    w = torch.full((10,), float(rank))
    parallel.broadcast_([w], src=0)
    # the non-blocking form the trainer uses (issue, do independent work, wait): same SUM
    g3 = torch.from_numpy(np.random.RandomState(200 + rank).randn(333).astype(np.float32))
    wait = parallel.allreduce_sum_async_(g3)
    busy = float(torch.ones(16).sum())           # (work that does not touch the buffer)
    wait()
    assert busy == 16.0
    out[rank] = dict(bounds=(b, e), g1=g1.numpy().copy(), g2=g2.numpy().copy(), n=n, w=w.numpy().copy(),
                     g3=g3.numpy().copy())
    torch.distributed.destroy_process_group()


def test_two_rank_allreduce_and_sharding():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0]['bounds'] == (0, 4) and out[1]['bounds'] == (4, 8)
    towers1 = [np.random.RandomState(100 + r).randn(1000).astype(np.float32) for r in range(world)]
    want = towers1[0] + towers1[1]
    for r in range(world):
        assert out[r]['n'] == 2                                   # the zero-sized buffer is skipped (tfutil.py:329)
        assert np.array_equal(out[r]['g1'], want)                 # both ranks hold the same SUM
        assert np.array_equal(out[r]['w'], np.zeros(10, np.float32))
        assert np.array_equal(out[r]['g3'], sum(np.random.RandomState(200 + q).randn(333).astype(np.float32)
                                                 for q in range(world)))
    # the summed gradient x 1/N drives the same Adam step as the oracle's tower list
    w_ref = np.ones(1000, np.float32)
    st = O.AdamState(1000, 0.0, 0.99)
    assert O.optimizer_step(w_ref, towers1, st, 0.0015)
    g = out[0]['g1'] * np.float32(0.5)
    lr_t = np.float32(0.0015) * np.sqrt(np.float32(1) - np.float32(0.99))
    v = np.float32(0.01) * g * g
    w = np.ones(1000, np.float32) - lr_t * g / (np.sqrt(v) + np.float32(1e-8))
    assert np.allclose(w, w_ref, rtol=0, atol=1e-7)


def test_shard_bounds_errors():
    from texturemixer_b200 import parallel
    assert parallel.shard_bounds(32, 8, 3) == (12, 16)
    try:
        parallel.shard_bounds(10, 4, 0)
        assert False
    except ValueError:
        pass
