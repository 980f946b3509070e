"""CPU checks of the training path's host logic: the product's orchestration (hmvit_b200/training.py) is run
with tests/emul_ops.py -- a torch restatement of every kernel, exact fp32 -- and its gradients are compared
with torch.autograd through the CPU oracle (which is pinned to the reference).  This validates the backward
algebra (folded-weight gradients, LayerNorm / GELU adjoints, recomputation, dead-query elimination, the
pull-back through the folding) without a GPU; the CUDA kernels themselves are checked op by op against the same
emulation in the GPU suite."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import hmvit_loader  # noqa: E402
from oracle import hmvit_oracle as O  # noqa: E402
import emul_ops  # noqa: E402


def _setup(B, L, H, W, record_len, seed, mode=None):
    pkg = hmvit_loader.load()
    cfg = O.default_config()
    cfg["hetero_fusion_block"]["drop_out"] = 0.0
    P = O.synth_state_dict(cfg, 0)
    net = pkg.HeteroFusion(cfg)
    net.load_state_dict(P, strict=True)
    x, T, m, rl, mask = O.synth_inputs(B, L, 256, H, W, record_len, seed=seed, tx=10, ty=5, mode=mode)
    return pkg, cfg, P, net, (x, T, m, rl, mask)


def _oracle_grads(cfg, P, inp, g_out, block_only=False):
    x, T, m, rl, mask = inp
    Pg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    xg = x.clone().requires_grad_(True)
    if block_only:
        y = O.fusion_block(xg, T, m.to(torch.int64), rl, mask, Pg, cfg["hetero_fusion_block"])
        valid = (torch.arange(x.shape[1])[None, :] < rl[:, None]).view(x.shape[0], x.shape[1], 1, 1, 1)
        loss = (torch.where(valid, y, torch.zeros(())) * g_out).sum()
    else:
        y = O.hetero_fusion(xg, T, m, rl, mask, Pg, cfg)
        loss = (y * g_out).sum()
    leaves = [xg] + [v for v in Pg.values() if v.is_floating_point()]
    names = ["x"] + [k for k, v in Pg.items() if v.is_floating_point()]
    gs = torch.autograd.grad(loss, leaves, allow_unused=True)
    return y.detach(), dict(zip(names, gs))


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("skip_dead", [True, False])
def test_fusion_gradients_match_oracle_autograd(skip_dead):
    pkg, cfg, P, net, inp = _setup(2, 3, 16, 24, [3, 2], seed=11)
    x, T, m, rl, mask = inp
    g_out = torch.randn(2, 256, 16, 24, generator=torch.Generator().manual_seed(3))
    y_ref, g_ref = _oracle_grads(cfg, P, inp, g_out)
    xg = x.clone().requires_grad_(True)
    y = pkg.training.fusion_train(emul_ops, net.hetero_fusion_block, net, xg, T, m, rl, mask, num_iters=net.num_iters,
                                  skip_dead=skip_dead)
    assert _rel(y.detach(), y_ref) < 2e-5
    (y * g_out).sum().backward()
    assert _rel(xg.grad, g_ref["x"]) < 2e-4, _rel(xg.grad, g_ref["x"])
    used = 0
    for name, p in net.named_parameters():
        ref = g_ref[name]
        if "aggregate_fc" in name:
            assert ref is None and (p.grad is None or float(p.grad.abs().max()) == 0.0)
            continue
        assert ref is not None, name
        if float(ref.norm()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) < 1e-6, name
            continue
        assert p.grad is not None, name
        err = _rel(p.grad, ref)
        assert err < 5e-4, (name, err)
        used += 1
    assert used > 40


def test_block_gradients_match_oracle_autograd():
    pkg, cfg, P, net, inp = _setup(1, 3, 16, 16, [2], seed=5, mode=[[1, 0, 0]])
    x, T, m, rl, mask = inp
    g_out = torch.randn(1, 3, 256, 16, 16, generator=torch.Generator().manual_seed(4))
    y_ref, g_ref = _oracle_grads(cfg, P, inp, g_out, block_only=True)
    blk = net.hetero_fusion_block
    xg = x.clone().requires_grad_(True)
    y = pkg.training.fusion_train(emul_ops, blk, None, xg, T, m, rl, mask, num_iters=1)
    assert _rel(y[:, :2].detach(), y_ref[:, :2]) < 2e-5
    assert torch.equal(y[:, 2].detach(), x[:, 2])                      # padded slot passes through
    valid = (torch.arange(3)[None, :] < rl[:, None]).view(1, 3, 1, 1, 1)
    (y * torch.where(valid, g_out, torch.zeros(()))).sum().backward()
    assert _rel(xg.grad[:, :2], g_ref["x"][:, :2]) < 2e-4
    for name, p in blk.named_parameters():
        ref = g_ref["hetero_fusion_block." + name]
        if ref is None or float(ref.norm()) == 0.0:
            continue
        assert _rel(p.grad, ref) < 5e-4, name


def test_fold_matches_inference_packing():
    """training.fold_stage must produce exactly the operands the inference path packs."""
    pkg, cfg, P, net, _ = _setup(1, 2, 8, 8, [2], seed=1)
    blk = net.hetero_fusion_block
    with torch.no_grad():
        packed = blk.packed()
        for kind in ("window", "grid"):
            F = pkg.training.fold_stage(blk, kind)
            pk = pkg.training.kernel_pack_stage(F, torch.bfloat16)
            ref = packed[kind]
            for key in ("wqkv0", "wqkv1", "wa0", "wa1", "w1_0", "w1_1", "w2_0", "w2_1", "bk", "bv", "ba", "b1", "b2", "bias_table"):
                assert torch.equal(pk[key], ref[key]), (kind, key)
            assert torch.allclose(pk["bcat"], ref["bqkv"], atol=1e-6), kind


def test_training_mode_dropout_is_rejected_and_cpu_tensors_raise():
    pkg, cfg, P, net, inp = _setup(1, 2, 8, 8, [2], seed=1)
    x, T, m, rl, mask = inp
    with pytest.raises(ValueError):
        net.eval()(x.requires_grad_(True), T, m, rl, mask)               # CPU tensors: no fallback


def test_gradients_match_reference_autograd_golden():
    """The same case against autograd through the REFERENCE itself (tests/golden/grads_c256.npz, made by
    tests/golden/make_golden_grads.py): dL/dx (strided sample + norm), every parameter gradient (norm + 64-value
    sample), and the set of parameters that receive no gradient -- for the oracle's autograd AND for the product's
    backward orchestration (training.fusion_train over the exact-fp32 kernel restatements)."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "grads_c256.npz"))
    pkg, cfg, P, net, inp = _setup(2, 3, 16, 24, [3, 2], seed=11)
    x, T, m, rl, mask = inp
    g_out = torch.randn(2, 256, 16, 24, generator=torch.Generator().manual_seed(3))
    names = [str(n) for n in g["names"]]
    nograd = {str(n) for n in g["nograd"]}
    dx_ref = torch.from_numpy(g["dx_sample"])
    samples = torch.from_numpy(g["samples"])

    def compare(dx, grads, tol_x, tol_p):
        assert _rel(dx[:, :, ::8, ::2, ::2], dx_ref) < tol_x
        assert float(dx.double().norm()) == pytest.approx(float(g["dx_norm"][0]), rel=tol_x)
        for k, name in enumerate(names):
            gr = grads[name]
            assert gr is not None, name
            flat = gr.reshape(-1)
            idx = torch.linspace(0, flat.numel() - 1, 64).long()
            assert float(flat.double().norm()) == pytest.approx(float(g["norms"][k]), rel=tol_p), name
            assert float((flat[idx] - samples[k]).norm()) <= tol_p * float(g["norms"][k]), name
        for name in nograd:
            gr = grads.get(name)
            assert gr is None or float(gr.abs().max()) < 1e-6, name

    y_o, g_o = _oracle_grads(cfg, P, inp, g_out)
    assert float(y_o.double().norm()) == pytest.approx(float(g["y_norm"][0]), rel=1e-5)
    compare(g_o["x"], g_o, 1e-4, 2e-4)

    xg = x.clone().requires_grad_(True)
    y = pkg.training.fusion_train(emul_ops, net.hetero_fusion_block, net, xg, T, m, rl, mask, num_iters=net.num_iters,
                                  skip_dead=True)
    (y * g_out).sum().backward()
    compare(xg.grad, {n: p.grad for n, p in net.named_parameters()}, 2e-4, 5e-4)


def _oracle_drop_masks(B, L, H, W, n_stage, seed, p):
    """The train-mode Dropout masks of the product path (Philox streams of hmvit_b200/training.py), exported in the
    oracle's (B, L, H, W, C) layout: list over stages of {"att", "hid", "ffn"}."""
    N = H * W
    pkg = hmvit_loader.load()
    out = []
    for s in range(n_stage):
        d = {}
        for site, name in enumerate(("att", "hid", "ffn")):
            m = emul_ops.dropout_mask_cm(B * L, N, seed, pkg.training._drop_stream(s, site), p)      # (B*L, 256, N)
            d[name] = m.view(B, L, 256, H, W).permute(0, 1, 3, 4, 2).contiguous()
        out.append(d)
    return out


def test_dropout_training_matches_oracle_with_replayed_masks():
    """Train mode with the shipped yaml's drop_out = 0.1: the product orchestration (forward with the three Dropout sites per
    stage + backward that regenerates the masks) against torch.autograd through the oracle with the SAME masks replayed
    (hetero_fusion.py:66, base_transformer.py:186-190 sites)."""
    pkg, cfg, P, net, inp = _setup(2, 3, 16, 24, [3, 2], seed=13)
    x, T, m, rl, mask = inp
    p_drop, seed = 0.1, 123456789
    g_out = torch.randn(2, 256, 16, 24, generator=torch.Generator().manual_seed(5))
    masks = _oracle_drop_masks(2, 3, 16, 24, 2 * cfg["num_iters"], seed, p_drop)
    Pg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    xr = x.clone().requires_grad_(True)
    y_ref = O.hetero_fusion(xr, T, m, rl, mask, Pg, cfg, drop_masks=masks)
    names = [k for k, v in Pg.items() if v.is_floating_point()]
    gs = torch.autograd.grad((y_ref * g_out).sum(), [xr] + [Pg[k] for k in names], allow_unused=True)
    g_ref = dict(zip(["x"] + names, gs))
    y_eval = O.hetero_fusion(x, T, m, rl, mask, P, cfg)
    assert _rel(y_ref.detach(), y_eval) > 0.05                      # the masks really change the result

    xg = x.clone().requires_grad_(True)
    y = pkg.training.fusion_train(emul_ops, net.hetero_fusion_block, net, xg, T, m, rl, mask, num_iters=net.num_iters,
                                  skip_dead=True, drop_p=p_drop, seed=seed)
    assert _rel(y.detach(), y_ref.detach()) < 1e-4
    (y * g_out).sum().backward()
    assert _rel(xg.grad, g_ref["x"]) < 5e-4
    worst = 0.0
    for name, prm in net.named_parameters():
        ref = g_ref.get(name)
        if ref is None or "aggregate_fc" in name:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, name
            continue
        if float(ref.norm()) < 1e-6:
            continue
        worst = max(worst, _rel(prm.grad, ref))
    assert worst < 2e-3, worst


def test_dropout_mask_statistics_and_streams():
    """Keep rate of the Philox mask within 4 sigma of 1 - p, values in {0, 1/(1-p)}, independent streams."""
    p = 0.1
    a = emul_ops.dropout_mask_cm(2, 1024, 77, 0, p)
    b = emul_ops.dropout_mask_cm(2, 1024, 77, 1, p)
    c = emul_ops.dropout_mask_cm(2, 1024, 78, 0, p)
    n = a.numel()
    scale = 1.0 / (1.0 - float(torch.tensor(p, dtype=torch.float32)))
    vals = sorted(torch.unique(a).tolist())
    assert len(vals) == 2 and vals[0] == 0.0 and vals[1] == pytest.approx(scale)
    for m in (a, b, c):
        keep = float((m != 0).float().mean())
        assert abs(keep - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5
    for u, v in ((a, b), (a, c)):
        agree = float(((u != 0) == (v != 0)).float().mean())
        assert abs(agree - ((1 - p) ** 2 + p ** 2)) < 0.01              # independent masks agree with prob. 0.82
    assert torch.equal(a, emul_ops.dropout_mask_cm(2, 1024, 77, 0, p))  # pure function of (seed, stream, index)


def test_module_train_mode_dispatch_cpu():
    """The module surface in train mode no longer rejects drop_out > 0: it draws a Philox seed per forward (or uses
    block.dropout_seed); CPU tensors still raise (no fallback)."""
    pkg, cfg, P, net, inp = _setup(1, 2, 8, 8, [2], seed=1)
    blk = net.hetero_fusion_block
    blk.drop_out = 0.1
    net.train()
    p, seed = pkg.fusion._dropout_args(blk)
    assert p == pytest.approx(0.1) and blk.last_dropout_seed == seed
    blk.dropout_seed = 42
    assert pkg.fusion._dropout_args(blk) == (pytest.approx(0.1), 42)
    net.eval()
    assert pkg.fusion._dropout_args(blk) == (0.0, 0)
