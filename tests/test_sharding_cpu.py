"""CPU, world_size 2, gloo: the multi-rank host logic (cost-balanced utterance partition + the single
all-gather + reassembly in utterance order).  The per-rank solve is injected (the oracle) because
there is no GPU here; on the GPU box the same function runs the CUDA path (tests/test_sharding_gpu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import windows_set


def test_partition_is_balanced_and_complete():
    from nnmnkwii_b200.sharding import partition
    rng = np.random.default_rng(0)
    lens = rng.integers(200, 2001, size=8192)  # BASELINE.json configs[4] lengths
    for world in (1, 2, 4, 8):
        parts = partition(lens, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(lens)))
        loads = np.array([lens[p].sum() for p in parts])
        assert loads.max() - loads.min() <= lens.max()
        assert loads.max() / loads.mean() < 1.001


def _oracle_solve(m, v, w, lens, lay):
    out = np.zeros((m.shape[0], lay.D_out), dtype=m.dtype)
    off = np.concatenate([[0], np.cumsum(lens)])
    sd = lay.D_out
    for u in range(len(lens)):
        a, b = off[u], off[u + 1]
        out[a:b, :sd] = oracle.mlpg(m[a:b], v if np.asarray(v).ndim == 1 else v[a:b], w)
    return out


def _worker(rank, world, port, lens, m, v, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nnmnkwii_b200.sharding import mlpg_batch_sharded
    y = mlpg_batch_sharded(m, v, windows_set()[2], lens, solve_fn=_oracle_solve)
    ret[rank] = y.numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    rng = np.random.default_rng(1)
    lens = rng.integers(3, 40, size=11)
    n = int(lens.sum())
    m = rng.random((n, 6)).astype(np.float32)
    v = (rng.random((n, 6)) + 0.1).astype(np.float32)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, lens, m, v, ret), nprocs=2, join=True)
    from nnmnkwii_b200 import paramgen as G
    ref = _oracle_solve(m, v, windows_set()[2], lens, G.StreamLayout.single(6, 3))
    assert np.array_equal(ret[0], ref) and np.array_equal(ret[1], ref)
    _ = (pytest, torch)
