"""GPU, needs >= 2 devices: the sharded MLPG path end to end (torchrun-style spawn, NCCL all-gather)
against the single-GPU result.  Skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import windows_set

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, lens, m, v, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from nnmnkwii_b200 import paramgen as G
    from nnmnkwii_b200.sharding import mlpg_batch_sharded
    y = mlpg_batch_sharded(m, v, windows_set()[2], lens, layout=G.merlin_layout())
    ret[rank] = y.cpu().numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from nnmnkwii_b200 import paramgen as G
    rng = np.random.default_rng(5)
    lens = rng.integers(20, 200, size=17)
    n = int(lens.sum())
    m = rng.random((n, 187), dtype=np.float32)
    v = rng.random((n, 187), dtype=np.float32) + 0.1
    ref = G.mlpg_batch(m, v, windows_set()[2], lengths=lens, layout=G.merlin_layout())
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, lens, m, v, ret), nprocs=2, join=True)
    assert np.array_equal(ret[0], ref) and np.array_equal(ret[1], ref)
