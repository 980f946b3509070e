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
