"""Generate the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Every case is: seeded synthetic inputs/parameters (oracle.hmvit_oracle.synth_*,
regenerated identically by the tests) -> the UNMODIFIED reference module ->
outputs saved as .npz.  Input checksums are stored so that a drift in the
synthetic generators is detected instead of silently invalidating the vectors.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from oracle import hmvit_oracle as O  # noqa: E402

CASES_FUSION = {
    # name: (input_dim, B, L, H, W, record_len, seed, tx, ty)
    "fusion_c64": (64, 2, 3, 16, 16, [3, 2], 11, 8.0, 6.0),
    "fusion_c256": (256, 2, 3, 16, 24, [2, 3], 12, 10.0, 6.0),
}


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def gen_fusion(R):
    for name, (C, B, L, H, W, rl, seed, tx, ty) in CASES_FUSION.items():
        cfg = O.default_config(input_dim=C)
        P = O.synth_state_dict(cfg, seed)
        x, T, mode, record_len, mask = O.synth_inputs(B, L, C, H, W, rl, seed + 100, tx=tx, ty=ty)
        ref = R.HeteroFusion(cfg).eval()
        ref.load_state_dict(P, strict=True)
        with torch.no_grad():
            out = ref(x.clone(), T.clone(), mode.clone(), record_len.clone(), mask.clone())
            blk = ref.hetero_fusion_block(x.clone(), T.clone(), mode.clone(), record_len.clone(), mask.clone())
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"),
                            meta=np.array([C, B, L, H, W, seed, tx, ty], dtype=np.float64),
                            record_len=np.array(rl), in_checksum=np.array([checksum(x), checksum(T)]),
                            p_checksum=np.array([sum(checksum(v) for v in P.values())]),
                            fused=out.numpy(), block=blk.numpy())
        print(name, tuple(out.shape), "ok")


def gen_attention(R):
    """HeteroAttention.forward on window-partitioned input with a random key mask, incl. a fully
    masked collaborator (hetero_fusion.py:187-277)."""
    C, b, l, X, Y, w = 256, 2, 3, 2, 2, 8
    g = torch.Generator().manual_seed(21)
    cfg = O.default_config(input_dim=C)
    P = O.synth_state_dict(cfg, 21)
    pfx = "hetero_fusion_block.grid_attention"
    att = R.HeteroAttention(C, 32, 0.1, 5, w).eval()
    att.load_state_dict({k[len(pfx) + 1:]: v for k, v in P.items() if k.startswith(pfx + ".")}, strict=True)
    x = torch.randn(b, l, X, Y, w, w, C, generator=g)
    mode = torch.tensor([[1, 0, 1], [0, 0, 1]], dtype=torch.int32)
    mask = (torch.rand(b, X, Y, w, w, 1, l, generator=g) > 0.3).float()
    mask[..., 0] = 1.0          # ego always visible
    mask[1, ..., 2] = 0.0       # one collaborator fully masked
    with torch.no_grad():
        out = att(x, mode, mask=mask)
    np.savez_compressed(os.path.join(HERE, "attention.npz"), x_checksum=np.array([checksum(x)]), mode=mode.numpy(),
                        mask=mask.numpy().astype(np.uint8), out=out.numpy())
    print("attention", tuple(out.shape), "ok")


def gen_warp_mask(R):
    """SpatialTransformation.forward and get_roi_and_cav_mask on random poses, 48x176."""
    B, L, C, H, W = 2, 4, 2, 48, 176
    x, T, mode, record_len, mask = O.synth_inputs(B, L, C, H, W, [4, 3], 31)
    st = R.SpatialTransformation({"voxel_size": [0.4, 0.4, 4], "downsample_rate": 4})
    warped, masks = [], []
    with torch.no_grad():
        for i in range(L):
            warped.append(st(x, T[:, :, i].clone()))
            masks.append(R.get_roi_and_cav_mask((B, L, H, W, C), mask, T[:, :, i].clone(), 0.4, 4))
    warped = torch.stack(warped, 2)                      # (B, L, L, C, H, W)  [b, src, tgt]
    masks = torch.stack(masks, -1)                       # (B, H, W, 1, L, L)  [..., src, tgt]
    np.savez_compressed(os.path.join(HERE, "warp_mask.npz"), in_checksum=np.array([checksum(x), checksum(T)]),
                        warped=warped.numpy(), mask_pair=np.packbits(masks.numpy().astype(np.uint8)),
                        mask_shape=np.array(masks.shape))
    print("warp_mask", tuple(warped.shape), "ok")


def gen_index(R):
    """token -> (window / grid) tables from the reference's einops patterns (hetero_fusion.py:384-389,
    427-431), the regroup mask (fuse_utils.py:8-61) and the relative position index buffer."""
    from einops import rearrange
    out = {}
    for (H, W) in ((48, 176), (128, 128), (96, 352), (16, 24)):
        idx = torch.arange(H * W).view(1, 1, 1, H, W)
        win = rearrange(idx, "b m d (x w1) (y w2) -> b m x y w1 w2 d", w1=8, w2=8).reshape(-1, 64)
        grd = rearrange(idx, "b m d (w1 x) (w2 y) -> b m x y w1 w2 d", w1=8, w2=8).reshape(-1, 64)
        out[f"window_{H}x{W}"] = win.numpy().astype(np.int32)
        out[f"grid_{H}x{W}"] = grd.numpy().astype(np.int32)
    att = R.HeteroAttention(256, 32, 0.1, 5, 8)
    out["relative_position_index"] = att.relative_position_index.numpy()
    dense = torch.randn(6, 2, 4, 4, generator=torch.Generator().manual_seed(5))
    feat, m = R.regroup(dense, torch.tensor([1, 3, 2]), 4)
    out["regroup_mask"] = m.numpy()
    out["regroup_feat_checksum"] = np.array([checksum(feat)])
    np.savez_compressed(os.path.join(HERE, "index.npz"), **out)
    print("index ok")


def gen_logits(R):
    """Detection-head logits of the reference pipeline: HeteroFusion -> HeteroDecoder (eval-mode BN,
    use_upsample=False) on the fusion_c256 case (hetero_decoder.py:42-74)."""
    C, B, L, H, W, rl, seed, tx, ty = CASES_FUSION["fusion_c256"]
    cfg = O.default_config(input_dim=C)
    P = O.synth_state_dict(cfg, seed)
    PD = O.synth_decoder_state_dict(seed + 1)
    x, T, mode, record_len, mask = O.synth_inputs(B, L, C, H, W, rl, seed + 100, tx=tx, ty=ty)
    ref = R.HeteroFusion(cfg).eval()
    ref.load_state_dict(P, strict=True)
    dec = R.HeteroDecoder({"input_dim": 256, "num_layer": 2, "num_ch_dec": [256, 256], "anchor_number": 2}).eval()
    dec.load_state_dict(PD, strict=True)
    with torch.no_grad():
        fused = ref(x.clone(), T.clone(), mode.clone(), record_len.clone(), mask.clone())
        psm, rm = dec(fused.unsqueeze(1), mode, use_upsample=False)
    np.savez_compressed(os.path.join(HERE, "logits_c256.npz"), psm=psm.numpy(), rm=rm.numpy(),
                        pd_checksum=np.array([sum(checksum(v.float()) for v in PD.values())]))
    print("logits", tuple(psm.shape), tuple(rm.shape), "ok")


if __name__ == "__main__":
    R = ref_import.load()
    torch.manual_seed(0)
    gen_index(R)
    gen_warp_mask(R)
    gen_attention(R)
    gen_fusion(R)
    gen_logits(R)
