"""Golden vectors of the detection post-processing from the UNMODIFIED reference:
VoxelPostprocessor.post_process (voxel_postprocessor.py:232-343) incl. box_utils.nms_rotated, run in the build container.

Modules the reference imports at module scope but that are absent here are replaced by stand-ins BEFORE the import:
  * shapely.geometry.Polygon  -> oracle.hmvit_postproc.ConvexPolygon (convex clipping in float64) -- the only ARITHMETIC that is
    not the reference's own: the intersection / union areas of the rotated NMS are parity-unpinned (see the oracle header);
  * opencood.utils.box_overlaps (an unbuilt cython extension), opencood.visualization.vis_utils (open3d / matplotlib):
    unused by post_process;
  * opencood.data_utils.datasets: only GT_RANGE is read from it (datasets/__init__.py:24).
Usage:  python tests/golden/make_golden_postproc.py   ->  tests/golden/postproc.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import hmvit_postproc as OP  # noqa: E402

REF = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    sys.path.insert(0, REF)
    _stub("shapely"); _stub("shapely.geometry", Polygon=OP.ConvexPolygon)
    _stub("opencood.utils.box_overlaps", bbox_overlaps=None)
    _stub("opencood.visualization"); _stub("opencood.visualization.vis_utils")
    _stub("opencood.data_utils.datasets", GT_RANGE=[-102.4, -102.4, -3, 102.4, 102.4, 1])
    for name in ("matplotlib", "matplotlib.pyplot", "cv2", "open3d"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    from opencood.data_utils.post_processor.voxel_postprocessor import VoxelPostprocessor
    return VoxelPostprocessor


def params(H, W, stride=4):
    rng = [-102.4, -102.4, -3, 102.4, 102.4, 1]
    vw = (rng[3] - rng[0]) / (W * stride)
    vh = (rng[4] - rng[1]) / (H * stride)
    return {"anchor_args": {"cav_lidar_range": rng, "l": 3.9, "w": 1.6, "h": 1.56, "r": [0, 90], "feature_stride": stride,
                            "num": 2, "W": W * stride, "H": H * stride, "vw": vw, "vh": vh},
            "target_args": {"pos_threshold": 0.6, "neg_threshold": 0.45, "score_threshold": 0.27},
            "order": "hwl", "max_num": 120, "nms_thresh": 0.15}


def synth_outputs(H, W, A, seed, density):
    """psm with a controlled fraction of anchors above the threshold, clustered so that the NMS has work to do; rm small."""
    g = torch.Generator().manual_seed(seed)
    psm = torch.full((1, A, H, W), -4.0)
    n_obj = max(1, int(density * H * W))
    ys = torch.randint(1, H - 1, (n_obj,), generator=g)
    xs = torch.randint(1, W - 1, (n_obj,), generator=g)
    for y, x in zip(ys.tolist(), xs.tolist()):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                for a in range(A):
                    psm[0, a, y + dy, x + dx] = torch.randn((), generator=g) * 1.5 + (1.0 if (dy, dx) == (0, 0) else -0.5)
    rm = torch.randn(1, 7 * A, H, W, generator=g) * 0.15
    rm[:, 2::7] = torch.randn(1, A, H, W, generator=g) * 0.05 - 0.1         # z near the anchor's
    return psm, rm


CASES = {"small": (16, 24, 11, 0.05), "bev": (48, 176, 12, 0.01)}


def main():
    VP = load_reference()
    out = {}
    for name, (H, W, seed, density) in CASES.items():
        P = params(H, W)
        pp = VP(P, train=False)
        anchors = pp.generate_anchor_box()
        assert np.array_equal(anchors, OP.generate_anchor_box(P)) and anchors.shape == (H, W, 2, 7)
        psm, rm = synth_outputs(H, W, 2, seed, density)
        th = 0.3
        T = torch.tensor([[np.cos(th), -np.sin(th), 0, 1.5], [np.sin(th), np.cos(th), 0, -2.0], [0, 0, 1, 0.1], [0, 0, 0, 1]], dtype=torch.float32)
        for tag, cav in (("proj", {"transformation_matrix": T, "anchor_box": torch.from_numpy(anchors)}),
                         ("noproj", {"transformation_matrix": T, "anchor_box": torch.from_numpy(anchors), "no_post_projection": True})):
            boxes, scores = pp.post_process({"ego": cav}, {"ego": {"psm": psm.clone(), "rm": rm.clone()}})
            ob, os_ = OP.post_process(psm, rm, torch.from_numpy(anchors), T if tag == "proj" else None, P)
            assert torch.equal(boxes, ob) and torch.equal(scores, os_), (name, tag)       # oracle == reference control flow + arithmetic
            out[f"{name}_{tag}_boxes"] = boxes.numpy()
            out[f"{name}_{tag}_scores"] = scores.numpy()
            print(name, tag, "boxes", tuple(boxes.shape))
        out[f"{name}_meta"] = np.array([H, W, seed, density], dtype=np.float64)
        out[f"{name}_T"] = T.numpy()
    np.savez_compressed(os.path.join(HERE, "postproc.npz"), **out)


if __name__ == "__main__":
    main()
