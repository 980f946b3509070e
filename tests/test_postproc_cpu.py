"""CPU tests of the detection post-processing: the oracle restatement against the golden vectors produced by
the UNMODIFIED reference post-processor (tests/golden/make_golden_postproc.py), the polygon stand-in against closed forms
(the one parity-unpinned piece), and the host mirror's anchor generation."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import hmvit_loader
from oracle import hmvit_postproc as OP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)
import make_golden_postproc as G  # noqa: E402  (the case definitions; the reference itself is not imported here)


def test_oracle_post_process_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "postproc.npz"))
    for name, (H, W, seed, density) in G.CASES.items():
        P = G.params(H, W)
        anchors = torch.from_numpy(OP.generate_anchor_box(P))
        psm, rm = G.synth_outputs(H, W, 2, seed, density)
        T = torch.from_numpy(g[f"{name}_T"])
        for tag in ("proj", "noproj"):
            boxes, scores = OP.post_process(psm, rm, anchors, T if tag == "proj" else None, P)
            assert np.array_equal(boxes.numpy(), g[f"{name}_{tag}_boxes"]), (name, tag)
            assert np.array_equal(scores.numpy(), g[f"{name}_{tag}_scores"]), (name, tag)
            assert boxes.shape[0] > 50                                    # the cases exercise the NMS
            assert bool((scores[:-1] >= scores[1:]).all())                # picked in score order


def test_convex_polygon_areas_closed_forms():
    sq = OP.ConvexPolygon([(0, 0), (2, 0), (2, 2), (0, 2)])
    assert sq.area == pytest.approx(4.0)
    assert sq.intersection(sq).area == pytest.approx(4.0) and sq.union(sq).area == pytest.approx(4.0)
    shifted = OP.ConvexPolygon([(1, 1), (3, 1), (3, 3), (1, 3)][::-1])            # clockwise on purpose
    assert sq.intersection(shifted).area == pytest.approx(1.0) and sq.union(shifted).area == pytest.approx(7.0)
    far = OP.ConvexPolygon([(5, 5), (6, 5), (6, 6), (5, 6)])
    assert sq.intersection(far).area == 0.0
    c, r = 1.0, math.sqrt(2.0)                                                   # the same square turned by 45 degrees: octagon
    dia = OP.ConvexPolygon([(c + r, c), (c, c + r), (c - r, c), (c, c - r)])
    assert dia.area == pytest.approx(4.0)
    assert sq.intersection(dia).area == pytest.approx(8.0 * (math.sqrt(2.0) - 1.0), rel=1e-12)
    # a 3.9 x 1.6 box against itself turned by 90 degrees: the common part is the 1.6 x 1.6 square
    a = OP.ConvexPolygon([(-1.95, -0.8), (1.95, -0.8), (1.95, 0.8), (-1.95, 0.8)])
    b = OP.ConvexPolygon([(-0.8, -1.95), (0.8, -1.95), (0.8, 1.95), (-0.8, 1.95)])
    assert a.intersection(b).area == pytest.approx(2.56) and a.union(b).area == pytest.approx(2 * 6.24 - 2.56)


def test_host_mirror_anchor_box_and_errors():
    pkg = hmvit_loader.load()
    P = G.params(48, 176)
    pp = pkg.VoxelPostprocessor(P, train=False)
    assert np.array_equal(pp.generate_anchor_box(), OP.generate_anchor_box(P))
    psm, rm = torch.zeros(1, 2, 48, 176), torch.zeros(1, 14, 48, 176)
    cav = {"anchor_box": pp.generate_anchor_box(), "transformation_matrix": torch.eye(4)}
    with pytest.raises(ValueError):
        pp.post_process({"ego": cav}, {"ego": {"psm": psm, "rm": rm}})            # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        pp.post_process({"ego": cav, "1": cav}, {"ego": {"psm": psm, "rm": rm}, "1": {"psm": psm, "rm": rm}})
