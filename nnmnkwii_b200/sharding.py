"""Utterance sharding across the GPUs of one box.

Every utterance (every DTW pair) is independent, so the data path has NO collective: the batch is
partitioned by cost (frames) with a longest-first greedy rule, each rank runs the single-GPU path on
its slice, and ONE all-gather at the end gives every rank the full result.  One process per GPU,
``torch.distributed`` (NCCL over NVLink on the GPUs; gloo in the CPU tests) is only the plumbing.
"""
import numpy as np


def partition(costs, world_size):
    """Greedy longest-processing-time partition.  Returns a list (one entry per rank) of index arrays,
    each sorted ascending; deterministic, identical on every rank."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    buckets = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        buckets[r].append(int(i))
        load[r] += costs[i]
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def _flat_slice(lengths, idx):
    off = np.concatenate([[0], np.cumsum(lengths)])
    if len(idx) == 0:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate([np.arange(off[i], off[i + 1]) for i in idx])


def all_gather_rows(local, counts, group=None):
    """All-gather a (n_local, D) tensor whose row count differs per rank (``counts[r]`` rows on rank
    r).  One collective: the local block is padded to the largest count."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    cap = int(max(counts)) if len(counts) else 0
    D = local.shape[1]
    padded = torch.zeros((max(cap, 1), D), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world, max(cap, 1), D), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out.view(-1), padded.view(-1), group=group)
    else:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        out = torch.stack(parts)
    return [out[r, : int(counts[r])] for r in range(world)]


def mlpg_batch_sharded(means, variances, windows, lengths, layout=None, group=None, solve_fn=None, device=None):
    """MLPG over a flat (sum_T, D) batch sharded by utterance over the ranks of ``group``.

    Every rank passes the SAME full inputs (NumPy) and gets the full ``(sum_T, D_out)`` result.
    ``solve_fn(means_local, variances_local, windows, lengths_local, layout)`` defaults to the
    single-GPU CUDA path (:func:`nnmnkwii_b200.paramgen.mlpg_batch` on device tensors).
    """
    import torch
    import torch.distributed as dist

    from . import paramgen as G

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lengths = np.asarray(lengths, dtype=np.int64)
    if layout is None:
        layout = G.StreamLayout.single(means.shape[1], len(windows))
    parts = partition(lengths, world)
    mine = parts[rank]
    rows = _flat_slice(lengths, mine)
    var1d = np.asarray(variances).ndim == 1
    m_loc = np.ascontiguousarray(means[rows])
    v_loc = variances if var1d else np.ascontiguousarray(variances[rows])
    if solve_fn is None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())

        def solve_fn(m, v, w, lens, lay):
            return G.mlpg_batch(torch.from_numpy(m).to(device), torch.from_numpy(np.asarray(v)).to(device), w,
                                lengths=lens, layout=lay)
    y_loc = solve_fn(m_loc, v_loc, windows, lengths[mine], layout)
    if not type(y_loc).__module__.startswith("torch"):
        y_loc = torch.from_numpy(np.ascontiguousarray(y_loc))
    counts = [int(lengths[p].sum()) for p in parts]
    gathered = all_gather_rows(y_loc, counts, group)  # the ONE collective
    out = torch.empty((int(lengths.sum()), layout.D_out), dtype=y_loc.dtype, device=y_loc.device)
    for r in range(world):
        out[torch.from_numpy(_flat_slice(lengths, parts[r])).to(out.device)] = gathered[r]
    return out
