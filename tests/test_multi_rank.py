"""world_size-2 gloo test of the N>1 host logic: frames are sharded by rank with no data-path
collective; timings are max-reduced and solved counts sum-reduced; the union of the shards'
results equals the single-rank results."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from defslam_b200 import shard, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.helpers import emu_solve_batched
    tmpl, frames = synthetic.make_config_frames("C1", nframes=n_frames)
    lo, hi = shard.shard_range(n_frames, rank, world)
    rc, outs = emu_solve_batched(frames[lo:hi])     # CPU stand-in for the per-rank kernel launch
    assert rc == 0
    ms, solved = shard.reduce_job_stats(10.0 + 5.0 * rank, hi - lo, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, [o.nodes for o in outs]))
    dist.barrier()
    if rank == 0:
        q.put((ms, solved, gathered))
    dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_rank():
    n_frames, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms, solved, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 15.0 and solved == n_frames          # max over ranks, sum over ranks
    from tests.helpers import emu_solve_batched
    tmpl, frames = synthetic.make_config_frames("C1", nframes=n_frames)
    rc, single = emu_solve_batched(frames)
    got = {}
    for lo, nodes in gathered:
        for i, nd in enumerate(nodes):
            got[lo + i] = nd
    assert sorted(got) == list(range(n_frames))
    for i in range(n_frames):
        assert np.array_equal(got[i], single[i].nodes)


def _nrsfm_worker(rank, world, port, q):
    """keyframe pairs (Schwarp fits) and keyframes (SfN solves) are independent units: same sharding,
    no data-path collective"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from defslam_b200 import nrsfm
    from tests.emu import build
    api = nrsfm.Api(C.CDLL(build.build()), "emu_")     # CPU stand-in for the per-rank kernel launch
    win = nrsfm.make_window(31, n_keypoints=200, n_views=3)
    cases = nrsfm.schwarp_cases(win)
    lo, hi = shard.shard_range(len(cases), rank, world)
    xs = [api.schwarp_fit(c).x for c in cases[lo:hi]]
    ms, solved = shard.reduce_job_stats(3.0 + rank, hi - lo, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, xs))
    dist.barrier()
    if rank == 0:
        q.put((ms, solved, gathered))
    dist.destroy_process_group()


def test_two_rank_sharded_schwarp_fits_equal_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nrsfm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms, solved, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 4.0 and solved == 3
    import ctypes as C
    from defslam_b200 import nrsfm
    from tests.emu import build
    api = nrsfm.Api(C.CDLL(build.build()), "emu_")
    win = nrsfm.make_window(31, n_keypoints=200, n_views=3)
    single = [api.schwarp_fit(c).x for c in nrsfm.schwarp_cases(win)]
    got = {}
    for lo, xs in gathered:
        for i, x in enumerate(xs):
            got[lo + i] = x
    assert sorted(got) == [0, 1, 2]
    for i in range(3):
        assert np.array_equal(got[i], single[i])
