"""CPU model of the inter-warp software pipeline of the fused exact-DTW kernel (csrc/nnk_dtw.cu,
dtw_fused_kernel): row groups of 32 stream over the columns, lane 31 of group g publishes its row into a
DTW_CW-column ring with a progress counter, lane 0 of group g + 1 reads it; both sides only synchronise at
checkpoints every DTW_POLL steps.  Under random warp scheduling the model checks the three properties the
kernel relies on: no deadlock, no column read before it was written, no ring slot overwritten before the
consumer has read it (columns c and c - 1 are both read)."""
import random

CW, POLL, LANES = 128, 8, 32


def run(n_groups, n_warps, Ty, seed):
    rng = random.Random(seed)
    nsteps = Ty + LANES - 1
    prog = [0] * (n_groups + 1)          # published: columns written by group g
    cons = [0] * (n_groups + 1)          # published: steps completed by group g (as a consumer)
    ring = [dict() for _ in range(n_groups)]   # ring[g][slot] = column currently stored (written by group g)
    reads_done = [0] * (n_groups + 1)    # steps actually executed by group g (ground truth for the overwrite check)
    cur = [w if w < n_groups else None for w in range(n_warps)]   # group a warp is working on
    s0 = [0] * n_groups
    finished = 0
    idle_rounds = 0
    while finished < n_groups:
        w = rng.randrange(n_warps)
        g = cur[w]
        if g is None:
            idle_rounds += 1
            assert idle_rounds < 100000, "deadlock"
            continue
        feeds = g + 1 < n_groups
        # ---- checkpoint (dtw_fused_kernel: `if ((s0 & (DTW_POLL - 1)) == 0)`) ----
        if g > 0:
            cons[g] = s0[g]
        if feeds:
            prog[g] = max(0, s0[g] - 31)
        blocked = False
        if g > 0 and prog[g - 1] < min(Ty, s0[g] + POLL):
            blocked = True
        if feeds and (s0[g] + POLL - 1 - 31) - CW + 2 > cons[g + 1]:
            blocked = True
        if blocked:
            idle_rounds += 1
            assert idle_rounds < 200000, "deadlock: group %d stuck at step %d" % (g, s0[g])
            continue
        idle_rounds = 0
        # ---- POLL steps ----
        for s in range(s0[g], min(nsteps, s0[g] + POLL)):
            if g > 0:  # lane 0 reads boundary columns s and s - 1 of group g - 1
                for c in (s, s - 1):
                    if 0 <= c < Ty:
                        assert ring[g - 1].get(c % CW) == c, "group %d read column %d before / after its time" % (g, c)
            j = s - 31  # lane 31 writes column j of its row
            if feeds and 0 <= j < Ty:
                old = ring[g].get(j % CW)
                if old is not None:  # the consumer must be past steps old and old + 1
                    assert reads_done[g + 1] > old + 1, "group %d overwrote column %d unread" % (g, old)
                ring[g][j % CW] = j
            reads_done[g] = s + 1
        s0[g] += POLL
        if s0[g] >= nsteps:
            if feeds:
                prog[g] = Ty
            if g > 0:
                cons[g] = nsteps + CW
            reads_done[g] = nsteps + CW
            finished += 1
            nxt = g + n_warps
            cur[w] = nxt if nxt < n_groups else None
    return True


def test_pipeline_is_deadlock_free_and_never_reads_or_overwrites_early():
    for seed in range(40):
        assert run(n_groups=7, n_warps=3, Ty=300, seed=seed)        # several passes per warp, ring wraps twice
    assert run(n_groups=29, n_warps=16, Ty=900, seed=1)              # configs[3] geometry
    assert run(n_groups=2, n_warps=16, Ty=40, seed=2)                # shorter than one ring
    for seed in range(10):
        assert run(n_groups=9, n_warps=2, Ty=200, seed=seed)         # the minimum: producer and consumer are never the same warp
