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


# ---- config 4: data-parallel training step, gradients all-reduced in one flat bucket (world size 2, gloo, CPU) ----
def _train_worker(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import emul_ops
    from oracle import hmvit_oracle as O
    pkg = hmvit_loader.load()
    cfg = O.default_config()
    cfg["hetero_fusion_block"]["drop_out"] = 0.0
    net = pkg.HeteroFusion(cfg)
    net.load_state_dict(O.synth_state_dict(cfg, 0))
    # rank 0: LiDAR ego + camera collaborator; rank 1: all-camera scene -> the LiDAR weights are unused on rank 1
    mode = [[1, 0]] if rank == 0 else [[0, 0]]
    x, T, m, rl, mask = O.synth_inputs(1, 2, 256, 8, 16, [2], seed=50 + rank, tx=4, ty=2, mode=mode)
    y = pkg.training.fusion_train(emul_ops, net.hetero_fusion_block, net, x, T, m, rl, mask, num_iters=net.num_iters)
    y.square().mean().backward()
    local = {n: (None if p.grad is None else p.grad.clone()) for n, p in net.named_parameters()}
    bucket = pkg.FlatGradAllReduce(net)
    bucket.allreduce(dist)
    after = {n: (None if p.grad is None else p.grad.clone()) for n, p in net.named_parameters()}
    # second step: the gradients are views of the bucket now (no pack / unpack): zero, backward, reduce again
    attached = all(p.grad is None or p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    bucket.zero_grad()
    y = pkg.training.fusion_train(emul_ops, net.hetero_fusion_block, net, x, T, m, rl, mask, num_iters=net.num_iters)
    y.square().mean().backward()
    still = all(p.grad is None or p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    bucket.allreduce(dist)
    step2_bad = [n for n, p in net.named_parameters()
                 if (after[n] is None) != (p.grad is None) or (p.grad is not None and not torch.allclose(p.grad, after[n], rtol=1e-5, atol=1e-8))]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    if rank == 0:
        # compare here: tensors do not survive the queue once the worker has exited
        n_none, bad = 0, []
        for name, g in after.items():
            parts = [r[name] for r in gathered]
            if all(q is None for q in parts):
                if g is not None:
                    bad.append(name + ": should stay None")
                n_none += 1
                continue
            ref = sum(q for q in parts if q is not None) / world
            if g is None or not torch.allclose(g, ref, rtol=1e-6, atol=1e-8):
                bad.append(name)
        out.put((n_none, bad + step2_bad, bucket.nbytes, len(after), attached and still))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_train_worker, args=(r, world, 29741, out)) for r in range(world)]
    for p in procs:
        p.start()
    n_none, bad, nbytes, n_params, views_ok = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert 10_000_000 < nbytes < 10_100_000                    # 2.5 M fp32 parameters in ONE bucket
    assert views_ok                                            # .grad of every used parameter is a view into the bucket
    assert not bad, bad                                        # every gradient == mean over ranks
    assert n_none == 8 and n_params > 80                       # aggregate_fc (unused on every rank) stays None
