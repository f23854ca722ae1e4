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


def _oracle_bucket(windows):
    """Stand-in for the per-rank CUDA solve (there is no GPU here): the oracle, writing each utterance
    of the bucket into the rank's slot of the gather buffer exactly where the kernel would."""
    def solve(batch, b, windows_c, chains, n_chain, status):
        meta = batch.meta[b]
        for k in range(meta["n_utt"]):
            a, o, n = int(meta["utt_off"][k]), int(meta["out_off"][k]), int(meta["utt_len"][k])
            y = oracle.mlpg(batch.means[a:a + n].numpy(), batch.variances[a:a + n].numpy(), windows)
            batch.result[o:o + n] = torch.from_numpy(y)
    return solve


def test_plan_layout_is_complete_and_tight():
    from nnmnkwii_b200.sharding import ShardPlan
    rng = np.random.default_rng(0)
    lens = rng.integers(200, 2001, size=8192)  # BASELINE.json configs[4]
    for world in (1, 2, 4, 8):
        plan = ShardPlan(lens, world, n_buckets=4)
        seen = np.sort(np.concatenate([m for mem in plan.members for m in mem]))
        assert np.array_equal(seen, np.arange(len(lens)))
        # every utterance has its own rows inside its (bucket, rank) slot; slots do not overlap
        ends = plan.row_start + lens
        order = np.argsort(plan.row_start)
        assert (plan.row_start[order][1:] >= ends[order][:-1]).all()
        assert ends.max() <= plan.rows_total
        # dead rows: at most one utterance per bucket and rank
        assert plan.rows_total - lens.sum() <= plan.n_buckets * world * lens.max()
        loads = np.array([plan.frames_of_rank(r) for r in range(world)])
        assert loads.max() / loads.mean() < 1.002


def _worker(rank, world, port, lens, m, v, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nnmnkwii_b200 import sharding
    w = windows_set()[2]
    sharding._solve_bucket = _oracle_bucket(w)
    y = sharding.mlpg_batch_sharded(m, v, w, lens, device=torch.device("cpu"), n_buckets=3)
    res = sharding.mlpg_batch_sharded(m, v, w, lens, device=torch.device("cpu"), n_buckets=2, utterance_order=False)
    ret[rank] = (y.numpy(), np.concatenate([res.utterance(u).numpy() for u in range(len(lens))]))
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
    for r in (0, 1):
        assert np.array_equal(ret[r][0], ref) and np.array_equal(ret[r][1], ref)
    _ = pytest


def test_plan_handles_fewer_utterances_than_ranks_and_empty_utterances():
    from nnmnkwii_b200.sharding import ShardPlan
    for lens, world in (([5, 0, 3], 8), ([7], 4), ([0, 0], 2), ([], 2), ([4, 4, 4, 4, 4], 2)):
        lens = np.asarray(lens, dtype=np.int64)
        plan = ShardPlan(lens, world, n_buckets=4)
        members = [u for mem in plan.members for ids in mem for u in ids]
        assert sorted(members) == list(range(len(lens)))
        assert plan.rows_total >= int(lens.sum()) and plan.rows_local * world >= plan.rows_total - 0
        assert sum(plan.frames_of_rank(r) for r in range(world)) == int(lens.sum())
        # rows of different utterances never overlap
        spans = sorted((int(plan.row_start[u]), int(plan.row_start[u] + lens[u])) for u in range(len(lens)) if lens[u] > 0)
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
