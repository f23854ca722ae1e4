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
    y = mlpg_batch_sharded(m, v, windows_set()[2], lens, layout=G.merlin_layout())  # default transport: peer (IPC + copy engines)
    y2 = mlpg_batch_sharded(m, v, windows_set()[2], lens, layout=G.merlin_layout(), transport="nccl", n_buckets=2)
    res = mlpg_batch_sharded(m, v, windows_set()[2], lens, layout=G.merlin_layout(), utterance_order=False)
    y3 = np.concatenate([res.utterance(u).cpu().numpy() for u in range(len(lens))])
    assert np.array_equal(y.cpu().numpy(), y2.cpu().numpy()) and np.array_equal(y.cpu().numpy(), y3)
    # peer set-up failing on ONE rank (simulated): every rank agrees to fall back to the NCCL all-gather
    for mode in ("open", "alloc"):
        os.environ["NNK_PEER_FORCE_FAIL"] = mode
        y5 = mlpg_batch_sharded(m, v, windows_set()[2], lens, layout=G.merlin_layout(), transport="peer")
        assert np.array_equal(y.cpu().numpy(), y5.cpu().numpy())
    os.environ.pop("NNK_PEER_FORCE_FAIL")
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


def test_launch_follows_the_tensor_device_not_the_current_device():
    """a tensor on cuda:1 while cuda:0 is current must run on cuda:1 (C ABI DeviceGuard)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nnmnkwii_b200 import paramgen as G
    from nnmnkwii_b200.metrics import melcd
    rng = np.random.default_rng(6)
    m = rng.random((50, 9), dtype=np.float32)
    v = rng.random((50, 9), dtype=np.float32) + 0.1
    torch.cuda.set_device(0)
    ref = G.mlpg(torch.from_numpy(m).to("cuda:0"), torch.from_numpy(v).to("cuda:0"), windows_set()[2]).cpu().numpy()
    y1 = G.mlpg(torch.from_numpy(m).to("cuda:1"), torch.from_numpy(v).to("cuda:1"), windows_set()[2])
    assert y1.device.index == 1 and torch.cuda.current_device() == 0
    assert np.array_equal(y1.cpu().numpy(), ref)
    side = torch.cuda.Stream(device=1)
    with torch.cuda.stream(side):  # a non-default stream of the other device
        y2 = G.mlpg(torch.from_numpy(m).to("cuda:1"), torch.from_numpy(v).to("cuda:1"), windows_set()[2])
    side.synchronize()
    assert np.array_equal(y2.cpu().numpy(), ref)
    a = torch.from_numpy(rng.random((4, 30, 5), dtype=np.float32))
    b = torch.from_numpy(rng.random((4, 30, 5), dtype=np.float32))
    assert abs(melcd(a.to("cuda:1"), b.to("cuda:1")) - melcd(a.to("cuda:0"), b.to("cuda:0"))) < 1e-12
