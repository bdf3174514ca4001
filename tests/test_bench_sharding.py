"""bench.py's N>1 logic on CPU: two gloo ranks, one clip per rank, whole-job time = max over ranks, audio = sum over ranks."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert bench.dist_env() == (rank, world, rank)
    wl = bench.rank_workload(rank)
    elapsed, n_audio = 0.4 + 0.1 * rank, 66240 + 320 * rank          # rank 1 is slower and makes a slightly longer clip
    dist.barrier()
    t, tot = bench.reduce_over_ranks(dist, elapsed, n_audio, "cpu")
    out.put((rank, wl["seed"], wl["prompt"], t, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, p0, t0, a0), (r1, s1, p1, t1, a1) = res
    assert (s0, s1) == (0, 1) and p0 != p1                              # independent clips: own seed, own prompt
    assert t0 == t1 == pytest.approx(0.5)                               # max over ranks
    assert a0 == a1 == 66240 + 66560                                    # whole-job audio
    # the metric the bench prints: audio seconds of all ranks / slowest rank's time
    assert a0 / 24000 / t0 == pytest.approx((66240 + 66560) / 24000 / 0.5)


def test_single_rank_passthrough():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.reduce_over_ranks(None, 0.25, 1000, "cpu") == (0.25, 1000.0)
