"""World-size-2 gloo test (CPU) of the multi-GPU host logic: scene sharding + max-over-ranks timing.
The data path itself has no collective (scenes are independent)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hmvit_loader


def _worker(rank, world, port, n_scenes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = hmvit_loader.load()
    lo, hi = pkg.scene_shard(n_scenes, world, rank)
    owned = torch.zeros(n_scenes, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)                                    # test-only collective: every scene owned exactly once
    t = pkg.max_over_ranks(10.0 + rank, dist)                 # the slowest rank defines the step time
    if rank == 0:
        out.put((owned.tolist(), t, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_scene_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world, n_scenes = 2, 17
    procs = [ctx.Process(target=_worker, args=(r, world, 29731, n_scenes, out)) for r in range(world)]
    for p in procs:
        p.start()
    owned, t, (lo, hi) = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert owned == [1] * n_scenes
    assert t == 11.0
    assert (lo, hi) == (0, 9)


def test_scene_shard_properties():
    pkg = hmvit_loader.load()
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [pkg.scene_shard(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
