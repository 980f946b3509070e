"""Golden ROI masks of the REFERENCE on adversarial poses (SURVEY 8c "mask" row): yaw a multiple of 90 deg with
(a) half-cell translations (0.8 m, 2.4 m at 1.6 m per cell): EVERY source coordinate is a rounding tie, the
    reference's outcome depends on its own fp32 round-off (normalise / inverse / affine_grid chain,
    torch_transformation_utils.py:108-134, 216-355);
(b) whole-cell translations (1.6 m, 4.8 m): no ties, must be bit-exact.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_ties.py
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_import  # noqa: E402

B, L, H, W = 1, 8, 48, 176


def poses(shifts):
    T = torch.eye(4).repeat(B, L, 1, 1)
    k = 0
    for yaw in (0.0, math.pi / 2, math.pi, -math.pi / 2):
        for tx in shifts:
            c, s = math.cos(yaw), math.sin(yaw)
            T[0, k, 0, 0], T[0, k, 0, 1], T[0, k, 1, 0], T[0, k, 1, 1] = c, -s, s, c
            T[0, k, 0, 3], T[0, k, 1, 3] = tx, -tx
            k += 1
    return T


def main():
    R = ref_import.load()
    cav = torch.ones(B, L, dtype=torch.int64)
    out = {}
    for name, shifts in (("half_cell", (0.8, 2.4)), ("whole_cell", (1.6, 4.8))):
        T = poses(shifts)
        m = R.get_roi_and_cav_mask((B, L, H, W, 1), cav, T.clone(), 0.4, 4)
        assert set(m.unique().tolist()) <= {0.0, 1.0}
        out[name] = np.packbits(m.numpy().astype(np.uint8).ravel())
        out[name + "_T"] = T.numpy()
        print(name, tuple(m.shape), int(m.sum()))
    np.savez_compressed(os.path.join(HERE, "mask_ties.npz"), shape=np.array([B, H, W, 1, L]), **out)


if __name__ == "__main__":
    main()
