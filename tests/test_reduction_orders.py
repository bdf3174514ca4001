"""CPU restatement of two reduction rewrites of the decode step (csrc/decode_kernels.cu) — the bit-exactness arguments, executable.

1. p2_scores reduces EIGHT dot products together with a transposed butterfly (stages xor 16, 8, 4 swap halves of the per-lane values,
   then xor 1 and xor 2 on the survivor).  Claim: for every task the result is bit-identical to lane_tree_reduce (common.cuh), the
   reference's GGML_F32x8_REDUCE order (ggml.c:1405-1422), and it ends up on the lanes the kernel publishes from.
2. block_layernorm combines the 16 warp partials with a 4-level xor butterfly started from partial[lane & 15].  Claim: every lane
   ends with the same bits (the bracket logic needs all threads to take the same decision).
IEEE float32 / float64 additions in numpy are the same operations as __fadd_rn / DADD.
"""
import numpy as np


def lane_tree_reduce(a):
    """a: [32] float32 lane partials -> [32] results (every lane), stage order xor 16, 8, 4, 1, 2."""
    a = a.astype(np.float32).copy()
    lanes = np.arange(32)
    for m in (16, 8, 4, 1, 2):
        a = (a + a[lanes ^ m]).astype(np.float32)
    return a


def transposed_butterfly(r):
    """r: [32 lanes][8 tasks] float32 -> (value per lane, task index per lane), as written in p2_scores."""
    r = r.astype(np.float32).copy()
    lanes = np.arange(32)
    u16, u8, u4 = (lanes & 16) != 0, (lanes & 8) != 0, (lanes & 4) != 0
    new = r.copy()
    for i in range(4):
        keep = np.where(u16, r[:, i + 4], r[:, i]); send = np.where(u16, r[:, i], r[:, i + 4])
        new[:, i] = (keep + send[lanes ^ 16]).astype(np.float32)
    r = new.copy()
    for i in range(2):
        keep = np.where(u8, r[:, i + 2], r[:, i]); send = np.where(u8, r[:, i], r[:, i + 2])
        new[:, i] = (keep + send[lanes ^ 8]).astype(np.float32)
    r = new.copy()
    keep = np.where(u4, r[:, 1], r[:, 0]); send = np.where(u4, r[:, 0], r[:, 1])
    v = (keep + send[lanes ^ 4]).astype(np.float32)
    v = (v + v[lanes ^ 1]).astype(np.float32)
    v = (v + v[lanes ^ 2]).astype(np.float32)
    mine = u16 * 4 + u8 * 2 + u4 * 1
    return v, mine


def test_transposed_butterfly_is_lane_tree_reduce_per_task():
    rng = np.random.default_rng(0)
    for trial in range(200):
        scale = np.float32(10.0 ** rng.integers(-6, 6))
        r = (rng.standard_normal((32, 8)) * scale).astype(np.float32)
        if trial % 7 == 0: r[rng.integers(0, 32), rng.integers(0, 8)] = np.float32(1e30)     # cancellation-prone cases
        v, mine = transposed_butterfly(r)
        for lane in range(32):
            ref = lane_tree_reduce(r[:, mine[lane]])
            assert v[lane].tobytes() == ref[lane].tobytes(), (trial, lane, mine[lane])
        # the publishing lanes (lane & 3 == 0) cover the eight tasks exactly once
        assert sorted(mine[np.arange(32) % 4 == 0].tolist()) == list(range(8))


def test_layernorm_partial_butterfly_gives_every_lane_the_same_bits():
    rng = np.random.default_rng(1)
    lanes = np.arange(32)
    for trial in range(200):
        part = (rng.standard_normal(16) * 10.0 ** rng.integers(-8, 8)).astype(np.float64)
        q = part[lanes & 15].copy()
        for o in (8, 4, 2, 1):
            q = q + q[lanes ^ o]
        assert len({x.tobytes() for x in q}) == 1
        # and it is the pairwise tree the previous per-thread version computed: ((p0+p8)+(p4+p12)) + ...
        t = part.copy()
        for st in (8, 4, 2, 1):
            t[:st] = t[:st] + t[st:2 * st]
        assert q[0].tobytes() == t[0].tobytes()
