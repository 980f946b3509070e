"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the detection post-processing of XHwind/HM-ViT, in the
reference's order of operations; only tests/ may import it.

    post_process            opencood/data_utils/post_processor/voxel_postprocessor.py:232-343
    delta_to_boxes3d        voxel_postprocessor.py:345-397
    generate_anchor_box     voxel_postprocessor.py:24-70
    boxes_to_corners_3d, project_box3d, corner_to_standup_box_torch, get_mask_for_boxes_within_range_torch,
    nms_rotated, remove_large_pred_bbx, remove_bbx_abnormal_z
                            opencood/utils/box_utils.py:139-184, 258-296, 231-255, 326-357, 575-620, 722-772
    compute_iou, convert_format, rotate_points_along_z
                            opencood/utils/common_utils.py:120-158, 29-51

PINNING.  Everything except the polygon arithmetic is pinned to the reference's own code: tests/golden/make_golden_postproc.py
runs the UNMODIFIED VoxelPostprocessor.post_process / box_utils.nms_rotated of the reference (tests/golden/postproc.npz).
`shapely` (the reference's polygon library, GEOS underneath; requirements.txt, unpinned version) is absent from this image, so
the generator substitutes `ConvexPolygon` below for `shapely.geometry.Polygon`: **the intersection / union areas of the
rotated-NMS are parity-unpinned** (convex clipping in float64, checked against closed forms in
tests/test_postproc_cpu.py); the control flow around them (score order, top-1000, greedy suppression, filters, range mask)
is the reference's.
"""
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

GT_RANGE = [-102.4, -102.4, -3, 102.4, 102.4, 1]          # opencood/data_utils/datasets/__init__.py:24


# ----------------------------------------------------------------------------------------------
# convex polygon stand-in for shapely.geometry.Polygon (what compute_iou needs: intersection, union, area)
# ----------------------------------------------------------------------------------------------
class ConvexPolygon:
    def __init__(self, pts):
        self.pts = [(float(x), float(y)) for x, y in pts]

    @staticmethod
    def _signed_area(pts):
        s = 0.0
        for i in range(len(pts)):
            x0, y0 = pts[i]
            x1, y1 = pts[(i + 1) % len(pts)]
            s += x0 * y1 - x1 * y0
        return 0.5 * s

    @property
    def area(self):
        return abs(self._signed_area(self.pts)) if len(self.pts) >= 3 else 0.0

    def _ccw(self):
        return self.pts if self._signed_area(self.pts) >= 0.0 else self.pts[::-1]

    def intersection(self, other):
        poly = list(self._ccw())
        clip = other._ccw()
        for e in range(len(clip)):
            if not poly:
                break
            x1, y1 = clip[e]
            x2, y2 = clip[(e + 1) % len(clip)]
            out = []
            for i in range(len(poly)):
                px, py = poly[i]
                qx, qy = poly[(i + 1) % len(poly)]
                sp = (x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)
                sq = (x2 - x1) * (qy - y1) - (y2 - y1) * (qx - x1)
                if sp >= 0.0:
                    out.append((px, py))
                if (sp >= 0.0) != (sq >= 0.0):
                    t = sp / (sp - sq)
                    out.append((px + t * (qx - px), py + t * (qy - py)))
            poly = out
        return ConvexPolygon(poly)

    def union(self, other):
        return _AreaOnly(self.area + other.area - self.intersection(other).area)


class _AreaOnly:
    def __init__(self, a):
        self.area = a


# ----------------------------------------------------------------------------------------------
def generate_anchor_box(params: Dict) -> np.ndarray:
    """voxel_postprocessor.py:24-70: (H / stride, W / stride, num, 7) anchors, 'hwl': x, y, z, h, w, l, r."""
    aa = params['anchor_args']
    W, H = aa['W'], aa['H']
    r = [math.radians(e) for e in aa['r']]
    num = aa['num']
    assert num == len(r)
    vh, vw = aa['vh'], aa['vw']
    xr = [aa['cav_lidar_range'][0], aa['cav_lidar_range'][3]]
    yr = [aa['cav_lidar_range'][1], aa['cav_lidar_range'][4]]
    fs = aa.get('feature_stride', 2)
    x = np.linspace(xr[0] + vw, xr[1] - vw, W // fs)
    y = np.linspace(yr[0] + vh, yr[1] - vh, H // fs)
    cx, cy = np.meshgrid(x, y)
    cx = np.tile(cx[..., np.newaxis], num)
    cy = np.tile(cy[..., np.newaxis], num)
    cz = np.ones_like(cx) * -1.0
    w = np.ones_like(cx) * aa['w']
    l = np.ones_like(cx) * aa['l']
    h = np.ones_like(cx) * aa['h']
    r_ = np.ones_like(cx)
    for i in range(num):
        r_[..., i] = r[i]
    if params['order'] == 'hwl':
        return np.stack([cx, cy, cz, h, w, l, r_], axis=-1)
    if params['order'] == 'lhw':
        return np.stack([cx, cy, cz, l, h, w, r_], axis=-1)
    raise SystemExit('Unknown bbx order.')


def delta_to_boxes3d(deltas: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
    """voxel_postprocessor.py:345-397."""
    N = deltas.shape[0]
    deltas = deltas.permute(0, 2, 3, 1).contiguous().view(N, -1, 7)
    boxes3d = torch.zeros_like(deltas)
    ar = anchors.view(-1, 7).float()
    ad = torch.sqrt(ar[:, 4] ** 2 + ar[:, 5] ** 2)
    ad = ad.repeat(N, 2, 1).transpose(1, 2)
    ar = ar.repeat(N, 1, 1)
    boxes3d[..., [0, 1]] = torch.mul(deltas[..., [0, 1]], ad) + ar[..., [0, 1]]
    boxes3d[..., [2]] = torch.mul(deltas[..., [2]], ar[..., [3]]) + ar[..., [2]]
    boxes3d[..., [3, 4, 5]] = torch.exp(deltas[..., [3, 4, 5]]) * ar[..., [3, 4, 5]]
    boxes3d[..., 6] = deltas[..., 6] + ar[..., 6]
    return boxes3d


def boxes_to_corners_3d(boxes3d: torch.Tensor, order: str) -> torch.Tensor:
    """box_utils.py:139-184 (on a copy: the reference reorders its argument in place)."""
    boxes3d = boxes3d.clone()
    if order == 'hwl':
        boxes3d[:, 3:6] = boxes3d[:, [5, 4, 3]]
    template = boxes3d.new_tensor(([1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1],
                                   [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1])) / 2
    corners = boxes3d[:, None, 3:6].repeat(1, 8, 1) * template[None, :, :]
    ang = boxes3d[:, 6]
    cosa, sina = torch.cos(ang), torch.sin(ang)
    zeros, ones = ang.new_zeros(corners.shape[0]), ang.new_ones(corners.shape[0])
    rot = torch.stack((cosa, sina, zeros, -sina, cosa, zeros, zeros, zeros, ones), dim=1).view(-1, 3, 3).float()
    corners = torch.matmul(corners.view(-1, 8, 3).float(), rot).view(-1, 8, 3)
    corners += boxes3d[:, None, 0:3]
    return corners


def project_box3d(box3d: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """box_utils.py:258-296."""
    c = box3d.transpose(1, 2)
    c = torch.cat((c, torch.ones((c.shape[0], 1, 8))), dim=1)
    return torch.matmul(T, c)[:, :3, :].transpose(1, 2)


def remove_large_pred_bbx(b: torch.Tensor) -> torch.Tensor:
    """box_utils.py:722-751, quirks included (the 'z length' is measured on y and ANDed as a number)."""
    x_len = torch.max(b[:, :, 0], dim=1)[0] - torch.min(b[:, :, 0], dim=1)[0]
    y_len = torch.max(b[:, :, 1], dim=1)[0] - torch.min(b[:, :, 1], dim=1)[0]
    z_len = torch.max(b[:, :, 1], dim=1)[0] - torch.min(b[:, :, 1], dim=1)[0]
    index = torch.logical_and(x_len <= 6, y_len <= 6)
    return torch.logical_and(index, z_len)


def remove_bbx_abnormal_z(b: torch.Tensor) -> torch.Tensor:
    """box_utils.py:754-772."""
    return torch.logical_and(torch.min(b[:, :, 2], dim=1)[0] >= -3, torch.max(b[:, :, 2], dim=1)[0] <= 1)


def nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, threshold: float) -> np.ndarray:
    """box_utils.py:575-620 with ConvexPolygon for shapely's Polygon (see the module docstring)."""
    if boxes.shape[0] == 0:
        return np.array([], dtype=np.int32)
    b = boxes.cpu().detach().numpy()
    s = scores.cpu().detach().numpy()
    polygons = [ConvexPolygon([(bx[i, 0], bx[i, 1]) for i in range(4)]) for bx in b]
    ixs = s.argsort()[::-1][:1000]
    pick = []
    while len(ixs) > 0:
        i = ixs[0]
        pick.append(i)
        iou = np.array([polygons[i].intersection(polygons[j]).area / polygons[i].union(polygons[j]).area for j in ixs[1:]],
                       dtype=np.float32)
        remove_ixs = np.where(iou > threshold)[0] + 1
        ixs = np.delete(ixs, remove_ixs)
        ixs = np.delete(ixs, 0)
    return np.array(pick, dtype=np.int32)


def range_mask(boxes: torch.Tensor) -> torch.Tensor:
    """box_utils.py:326-357."""
    lo = torch.Tensor(GT_RANGE[:2]).reshape(1, 1, -1)
    hi = torch.Tensor(GT_RANGE[3:5]).reshape(1, 1, -1)
    return torch.all(torch.all(boxes[:, :, :2] >= lo, dim=-1) & torch.all(boxes[:, :, :2] <= hi, dim=-1), dim=-1)


def post_process(psm: torch.Tensor, rm: torch.Tensor, anchor_box: torch.Tensor, T: Optional[torch.Tensor], params: Dict
                 ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """voxel_postprocessor.py:232-343 for the single 'ego' entry of intermediate fusion (T = None: 'no_post_projection')."""
    prob = torch.sigmoid(psm.permute(0, 2, 3, 1)).reshape(1, -1)
    batch_box3d = delta_to_boxes3d(rm, anchor_box)
    mask = torch.gt(prob, params['target_args']['score_threshold']).view(1, -1)
    mask_reg = mask.unsqueeze(2).repeat(1, 1, 7)
    assert batch_box3d.shape[0] == 1
    boxes3d = torch.masked_select(batch_box3d[0], mask_reg[0]).view(-1, 7)
    scores = torch.masked_select(prob[0], mask[0])
    if len(boxes3d) == 0:
        return None, None
    corners = boxes_to_corners_3d(boxes3d, order=params['order'])
    proj = project_box3d(corners, T) if T is not None else corners
    keep = torch.logical_and(remove_large_pred_bbx(proj), remove_bbx_abnormal_z(proj))
    pred, scores = proj[keep], scores[keep]
    keep_index = nms_rotated(pred, scores, params['nms_thresh'])
    pred, scores = pred[keep_index], scores[keep_index]
    m = range_mask(pred)
    return pred[m, :, :], scores[m]
