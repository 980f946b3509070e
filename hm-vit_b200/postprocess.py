"""Detection post-processing on B200: mirror of
    VoxelPostprocessor     /root/reference/opencood/data_utils/post_processor/voxel_postprocessor.py:19-343
for what inference of the intermediate-fusion HM-ViT model uses: `generate_anchor_box()` (host, numpy, :24-70) and
`post_process(data_dict, output_dict)` (:232-343) -> (pred_box3d_tensor (M, 8, 3), scores (M,)) or (None, None).
The arithmetic of post_process runs in one C call (`hmvit_postprocess`, csrc/postproc.cuh): sigmoid + threshold, anchor
decoding, corners, projection, sanity filters, rotated NMS, range mask -- psm / rm never leave the device.  Label generation
for training (`generate_label`) and the visualisation helpers are not rebuilt.  No CPU fallback.
"""
import math
import sys

import numpy as np
import torch

from . import _lib, ops

GT_RANGE = [-102.4, -102.4, -3, 102.4, 102.4, 1]          # opencood/data_utils/datasets/__init__.py:24


class VoxelPostprocessor:
    def __init__(self, anchor_params, train=False):
        self.params = anchor_params
        self.train = train
        self.anchor_num = self.params['anchor_args']['num']

    def generate_anchor_box(self):
        aa = self.params['anchor_args']
        W, H = aa['W'], aa['H']
        l, w, h, r = aa['l'], aa['w'], aa['h'], aa['r']
        assert self.anchor_num == len(r)
        r = [math.radians(ele) for ele in r]
        vh, vw = aa['vh'], aa['vw']
        xrange = [aa['cav_lidar_range'][0], aa['cav_lidar_range'][3]]
        yrange = [aa['cav_lidar_range'][1], aa['cav_lidar_range'][4]]
        feature_stride = aa['feature_stride'] if 'feature_stride' in aa else 2
        x = np.linspace(xrange[0] + vw, xrange[1] - vw, W // feature_stride)
        y = np.linspace(yrange[0] + vh, yrange[1] - vh, H // feature_stride)
        cx, cy = np.meshgrid(x, y)
        cx = np.tile(cx[..., np.newaxis], self.anchor_num)
        cy = np.tile(cy[..., np.newaxis], self.anchor_num)
        cz = np.ones_like(cx) * -1.0
        w, l, h = np.ones_like(cx) * w, np.ones_like(cx) * l, np.ones_like(cx) * h
        r_ = np.ones_like(cx)
        for i in range(self.anchor_num):
            r_[..., i] = r[i]
        if self.params['order'] == 'hwl':
            return np.stack([cx, cy, cz, h, w, l, r_], axis=-1)
        if self.params['order'] == 'lhw':
            return np.stack([cx, cy, cz, l, h, w, r_], axis=-1)
        sys.exit('Unknown bbx order.')

    def post_process(self, data_dict, output_dict):
        keys = [k for k in data_dict if k in output_dict]
        if len(keys) != 1:
            raise NotImplementedError("hmvit_b200 post-processing covers intermediate fusion: exactly one cav ('ego') with "
                                      f"model outputs, got {keys}")
        cav = data_dict[keys[0]]
        psm, rm = output_dict[keys[0]]['psm'], output_dict[keys[0]]['rm']
        if not psm.is_cuda:
            raise ValueError("hmvit_b200 has no CPU path: psm / rm must be CUDA tensors")
        if psm.shape[0] != 1:
            raise AssertionError("during validation/testing, the batch size should be 1")      # (:279)
        if self.params['order'] not in ('hwl', 'lwh'):
            raise ValueError("order must be 'hwl' or 'lwh'")
        dev = psm.device
        A, H, W = psm.shape[1], psm.shape[2], psm.shape[3]
        anchors = torch.as_tensor(cav['anchor_box'], dtype=torch.float32).to(dev).contiguous()
        if tuple(anchors.shape) != (H, W, A, 7):
            raise ValueError(f"anchor_box {tuple(anchors.shape)} does not match psm {tuple(psm.shape)}")
        tmat = None
        if 'no_post_projection' not in cav:
            tmat = torch.as_tensor(cav['transformation_matrix'], dtype=torch.float32).to(dev).contiguous()
            assert tmat.shape == (4, 4)
        boxes, scores, meta = ops.postprocess(psm=psm.float().contiguous(), rm=rm.float().contiguous(), anchor_box=anchors,
                                                       transformation_matrix=tmat, order_hwl=self.params['order'] == 'hwl',
                                                       score_threshold=self.params['target_args']['score_threshold'],
                                                       nms_thresh=self.params['nms_thresh'], gt_range=GT_RANGE)
        n, st, n_cand = meta.tolist()                           # the one host read: how many boxes survived
        if st != 0:
            raise _lib.HmvitError("postprocess: more than 16384 anchors passed the score threshold and the sanity filters")
        if n_cand == 0:
            # (:313-314 returns (None, None) when no anchor passes the score threshold; with candidates that are all removed
            # by the sanity filters the reference fails in nms_rotated's empty-array indexing -- empty results here)
            return None, None
        return boxes[:n], scores[:n]
