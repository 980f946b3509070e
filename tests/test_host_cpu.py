"""CPU-side tests of the host logic and of the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

import hmvit_loader
from oracle import hmvit_emul as E
from oracle import hmvit_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    pkg = hmvit_loader.load()
    hdr = open(os.path.join(ROOT, "include", "hmvit_b200.h")).read()
    declared = set(re.findall(r"\b(hmvit_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(pkg._lib.EXPORTS)
    assert os.path.exists(pkg._lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert pkg._lib.load().hmvit_abi_version() == pkg._lib.ABI_VERSION == 9


def test_state_dict_keys_match_reference_spec():
    pkg = hmvit_loader.load()
    cfg = O.default_config()
    net = pkg.HeteroFusion(cfg)
    spec = O.state_dict_spec(cfg)
    sd = net.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    for k, shape in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sd["hetero_fusion_block.window_attention.relative_position_index"].dtype == torch.int64
    assert torch.equal(sd["hetero_fusion_block.grid_attention.relative_position_index"], O.relative_position_index(8))
    net.load_state_dict(O.synth_state_dict(cfg, 1), strict=True)


def test_constructor_rejects_unsupported_dims_and_cpu_input():
    pkg = hmvit_loader.load()
    with pytest.raises(ValueError):
        pkg.HeteroFusion(O.default_config(input_dim=128))
    net = pkg.HeteroFusion(O.default_config()).eval()
    x, T, mode, rl, mask = O.synth_inputs(1, 2, 256, 8, 8, [2], 0)
    with torch.no_grad(), pytest.raises(ValueError):
        net(x, T, mode, rl, mask)            # CPU tensors: the product path has no CPU fallback


def test_packed_weights_match_emulation_folding():
    """Host-side weight folding (scale, relation_att / relation_msg) == the oracle-side folding."""
    pkg = hmvit_loader.load()
    cfg = O.default_config()
    P = O.synth_state_dict(cfg, 2)
    net = pkg.HeteroFusion(cfg)
    net.load_state_dict(P, strict=True)
    pk = net.hetero_fusion_block.packed()["grid"]
    ref = E.pack_attention_weights(P, "hetero_fusion_block.grid_attention")
    for t in (0, 1):
        g1 = P[f"hetero_fusion_block.grid_norm.net.{t}.weight"]
        b1 = P[f"hetero_fusion_block.grid_norm.net.{t}.bias"]
        w = pk[f"wqkv{t}"].float()
        exp = torch.cat([ref["wq"][t] * 1.4426950408889634, ref["wk"][0][t], ref["wk"][1][t], ref["wv"][0][t], ref["wv"][1][t]], 0)
        expb = torch.cat([ref["bq"][t] * 1.4426950408889634, torch.zeros(1024)]) + exp @ b1     # LayerNorm beta folded
        assert torch.allclose(w, exp * g1[None, :], rtol=2 ** -8, atol=1e-6)                     # bf16 storage, gamma folded
        assert torch.allclose(pk["bqkv"][t], expb, atol=1e-5)
        for te in (0, 1):
            assert torch.allclose(pk["bk"][te, t], ref["bk"][te][t], atol=1e-7)
            assert torch.allclose(pk["bv"][te, t], ref["bv"][te][t], atol=1e-7)
    # tf32 rounding keeps 10 mantissa bits
    w1 = pk["w1_0"]
    assert int((w1.view(torch.int32) & 0x1FFF).abs().max()) == 0
    g2 = P["hetero_fusion_block.grid_ffd.norm.net.0.weight"]
    assert float((w1 - P["hetero_fusion_block.grid_ffd.fn.net.0.0.weight"] * g2[None, :]).abs().max()) < 1e-3
    # fp16 copies for the fused chain kernel: same 11-bit significand as tf32 -> they agree with the tf32 copies to one
    # ulp of that format wherever fp16 is normal (|w| >= 2^-14), and no weight overflows
    for key in ("w1", "w2"):
        for t in (0, 1):
            h, f = pk[f"{key}h_{t}"], pk[f"{key}_{t}"]
            assert h.dtype == torch.float16 and h.shape == f.shape and bool(torch.isfinite(h).all())
            normal = f.abs() >= 2.0 ** -14
            assert float(((h.float() - f).abs() / f.abs().clamp_min(1e-30))[normal].max()) <= 2.0 ** -10


def test_regroup_matches_oracle():
    pkg = hmvit_loader.load()
    dense = torch.randn(6, 3, 4, 4)
    rl = torch.tensor([1, 3, 2])
    a, ma = pkg.regroup(dense, rl, 4)
    b, mb = O.regroup(dense, rl, 4)
    assert torch.equal(a, b) and torch.equal(ma, mb) and ma.dtype == torch.int64


def test_emulation_restructuring_is_exact():
    """project-then-warp + folded edge weights + dead-query elimination == the reference order of ops."""
    cfg = O.default_config()
    P = O.synth_state_dict(cfg, 0)
    x, T, mode, rl, mask = O.synth_inputs(2, 3, 256, 16, 16, [3, 2], seed=4, tx=8, ty=5)
    yo = O.hetero_fusion(x, T, mode, rl, mask, P, cfg)
    ye = E.hetero_fusion(x, T, mode, rl, mask, P, cfg)
    assert float((ye - yo).norm() / yo.norm()) < 5e-6


def test_launch_accounting():
    """hmvit_fusion_launch_count is pure host code: one key-record pass per forward + per stage QKV + attention + chain;
    the head is a launch of its own only without dead-query elimination (with it, it runs inside the last stage's chain
    launch).  The split cross-check form has two attention launches per stage and no record pass."""
    pkg = hmvit_loader.load()
    ops = pkg.ops
    assert ops.fusion_launch_count(2, False) == 13
    assert ops.fusion_launch_count(2, True, skip_dead=False) == 14
    assert ops.fusion_launch_count(2, True, skip_dead=True) == 13
    assert ops.fusion_launch_count(1, True) == 7
    assert ops.fusion_launch_count(2, True, attn_impl="split") == 16
    assert ops.fusion_launch_count(2, True, attn_impl="single") == 12



def test_bench_flop_count_matches_survey():
    """190.5 GFLOP per config-2 scene with all agents as queries; with last-stage dead-query
    elimination (Lv-1)/Lv of that stage's Q projection, QK^T, PV, O and FFN terms are subtracted."""
    import importlib.util
    import os
    sp = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    b = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(b)
    assert abs(b.scene_flops(5) / 1e9 - 190.455) < 1e-3
    N, C, Lv = 8448, 256, 5
    skipped = (Lv - 1) * N * (2 * C * C + 2 * (Lv * 64) * 2 * C + 2 * C * C + 4 * C * C)
    assert b.scene_flops(5, 2, True) == b.scene_flops(5) - skipped
    assert abs(b.scene_flops(2) / 1e9 - 64.2) < 0.1                      # config 1 figure of the survey


def test_decoder_state_dict_keys_and_batchnorm_folding():
    """HeteroDecoder mirror: state_dict keys / shapes of the reference (oracle spec, which loads strict=True into the
    reference in tests/golden/make_golden.py), and the BatchNorm folding the kernels consume is exact: convolutions with
    the folded weights reproduce the restatement of hetero_decoder.py:42-74 in fp32."""
    import torch.nn.functional as F
    pkg = hmvit_loader.load()
    params = {"input_dim": 256, "num_layer": 2, "num_ch_dec": [256, 256], "anchor_number": 2}
    dec = pkg.HeteroDecoder(params).eval()
    spec = O.decoder_state_spec()
    sd = dec.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    for k, shape in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    PD = O.synth_decoder_state_dict(3)
    dec.load_state_dict(PD, strict=True)
    pk = dec.packed()
    assert pk["conv_w"].shape == (4, 2, 9, 256, 256) and pk["conv_w"].dtype == torch.float16
    assert pk["conv_b"].shape == (4, 2, 256) and pk["head_w"].shape == (2, 16, 256) and pk["head_b"].shape == (2, 16)
    torch.manual_seed(0)
    x = torch.randn(2, 256, 8, 8)
    ego_mode = torch.tensor([1, 0])
    rp, rr = O.hetero_decoder(x, ego_mode, PD)
    for b in range(2):
        t = int(ego_mode[b])
        nd = dec.lidar_decoder if t == 1 else dec.camera_decoder
        w, bias = nd.folded()
        y = x[b:b + 1]
        for l in range(4):
            y = F.relu(F.conv2d(y, w[l].view(3, 3, 256, 256).permute(2, 3, 0, 1), bias[l], padding=1))
        out = F.conv2d(y, pk["head_w"][t].view(16, 256, 1, 1), pk["head_b"][t])
        assert float((out[:, :2] - rp[b:b + 1]).norm() / rp[b:b + 1].norm()) < 1e-5
        assert float((out[:, 2:] - rr[b:b + 1]).norm() / rr[b:b + 1].norm()) < 1e-5
    with pytest.raises(NotImplementedError):
        dec(x, torch.ones(2, 1), use_upsample=True)
    with pytest.raises(ValueError):
        dec(x, torch.ones(2, 1), use_upsample=False)           # CPU tensors: no fallback
    with pytest.raises(ValueError):
        pkg.HeteroDecoder({"input_dim": 128, "num_layer": 2, "num_ch_dec": [128, 128], "anchor_number": 2})


def test_model_glue_matches_reference_restatement():
    """combine_features / unpad_mode_encoding / extract_lidar_input of the model glue (vectorised) against the loop
    restatement of base_camera_lidar_intermediate.py:31-99; the config keys of the shipped yaml construct the model."""
    pkg = hmvit_loader.load()
    from importlib import import_module
    M = import_module("hmvit_b200.model")
    torch.manual_seed(3)
    mode = torch.tensor([[1, 0, 0, 1, 0], [0, 1, 1, 0, 0], [1, 1, 0, 0, 0]])
    rl = torch.tensor([4, 2, 5])
    mu = M.unpad_mode_encoding(mode, rl)
    assert torch.equal(mu, O.unpad_mode_encoding(mode, rl))
    n_cam, n_lid = int((mu == 0).sum()), int((mu == 1).sum())
    cam, lid = torch.randn(n_cam, 4, 2, 3), torch.randn(n_lid, 4, 2, 3)
    assert torch.equal(M.combine_features(cam, lid, mode, rl), O.combine_features(cam, lid, mode, rl))
    x = torch.randn(3, 5, 2)
    assert torch.equal(M.unpad_features(x, rl), torch.cat([x[i, :int(rl[i])] for i in range(3)]))
    with pytest.raises(ValueError):
        M.combine_features(cam[:1], lid, mode, rl)
    nv = 57
    pl = {"voxel_features": torch.randn(nv, 32, 4), "voxel_coords": torch.randint(0, 40, (nv, 4)),
          "voxel_num_points": torch.randint(1, 32, (nv,))}
    pl["voxel_coords"][:, 0] = torch.randint(0, int(rl.sum()), (nv,))
    got = M.BevformerPointPillarHetero.extract_lidar_input({"processed_lidar": pl}, mu)["processed_lidar"]
    want = O.extract_lidar_input(pl, mu)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    cfg = {"max_cav": 5, "compression": 0, "anchor_number": 2,
           "spatial_transform": O.default_config()["spatial_transform"], "hetero_fusion": O.default_config(),
           "hetero_decoder": {"input_dim": 256, "num_layer": 2, "num_ch_dec": [256, 256], "anchor_number": 2}}
    net = M.BevformerPointPillarHetero(cfg)
    keys = list(net.state_dict().keys())
    assert any(k.startswith("fusion_net.hetero_fusion_block.") for k in keys) and "decoder.camera_cls_head.weight" in keys
    assert "cls_head.weight" in keys and "reg_head.bias" in keys
    with pytest.raises(NotImplementedError):
        net({"mode": mode, "record_len": rl, "pairwise_t_matrix": None})      # no encoders passed in
