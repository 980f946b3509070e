"""Model glue of the HM-ViT detector around the fusion hot path: mirrors
    BevformerPointPillarHetero          /root/reference/opencood/models/bevformer_point_pillar_hetero.py:52-134
    BaseCameraLiDARIntermediate          /root/reference/opencood/models/base_camera_lidar_intermediate.py:4-99
      (unpad_mode_encoding :68-73, unpad_features :74-79, combine_features :81-99)
    regroup                              /root/reference/opencood/models/fuse_utils.py:8-61  (hm-vit_b200/fusion.py:regroup)
i.e. everything between the two BEV encoders and the output dict:
    per-agent camera / LiDAR BEV features -> combine (agent order) -> regroup (B, L, C, H, W) + mask -> HeteroFusion
    (hmvit_fusion_forward) -> HeteroDecoder (hmvit_decoder_forward) -> {'psm', 'rm'}.
The encoders are passed in (`camera_encoder`, `lidar_encoder`: any modules mapping the reference's `extract_*_input(batch)`
dicts to (n, 256, H, W) features -- `hm-vit_b200/encoders.py` has library-backed mirrors of PointPillar and the CVT camera branch
and `build_config3_model`; the shipped yaml's BEVFormer / mmcv branch is not rebuilt), or the caller hands the
features to `forward_features` directly.  `compression > 0` (NaiveCompressor) is not
used by the shipped yaml (`hypes_yaml/opcl/bevformer_point_pillar_hetero.yaml:83`) and raises.
"""
import torch
import torch.nn as nn

from .decoder import HeteroDecoder
from .fusion import HeteroFusion, regroup


def unpad_mode_encoding(mode, record_len):
    """(B, L) modality flags -> (sum record_len,) flags of the valid agents in scene order
    (base_camera_lidar_intermediate.py:68-73), without the per-scene Python loop."""
    L = mode.shape[1]
    valid = torch.arange(L, device=mode.device)[None, :] < record_len.to(mode.device)[:, None]
    return mode[valid]


def unpad_features(x, record_len):
    """(B, L, ...) -> (sum record_len, ...) (base_camera_lidar_intermediate.py:74-79)."""
    L = x.shape[1]
    valid = torch.arange(L, device=x.device)[None, :] < record_len.to(x.device)[:, None]
    return x[valid]


def combine_features(camera_feature, lidar_feature, mode, record_len):
    """Interleave the camera agents' and the LiDAR agents' features back into agent order
    (base_camera_lidar_intermediate.py:81-99): the i-th agent of modality m takes the next row of that modality's
    tensor.  One masked scatter per modality instead of a Python loop over agents."""
    if mode.dim() == 2:
        mode = unpad_mode_encoding(mode, record_len)
    if bool(((mode != 0) & (mode != 1)).any()):
        raise ValueError("Mode but be either 1 or 0")
    src = camera_feature if camera_feature is not None else lidar_feature
    out = src.new_empty((mode.shape[0],) + tuple(src.shape[1:]))
    cam, lid = mode == 0, mode == 1
    n_cam, n_lid = int(cam.sum()), int(lid.sum())
    if n_cam:
        if camera_feature is None or camera_feature.shape[0] != n_cam:
            raise ValueError(f"{n_cam} camera agents but camera features {None if camera_feature is None else tuple(camera_feature.shape)}")
        out[cam] = camera_feature
    if n_lid:
        if lidar_feature is None or lidar_feature.shape[0] != n_lid:
            raise ValueError(f"{n_lid} lidar agents but lidar features {None if lidar_feature is None else tuple(lidar_feature.shape)}")
        out[lid] = lidar_feature
    return out


class BevformerPointPillarHetero(nn.Module):
    """Same config keys as the reference model (`model.args` of the yaml): 'max_cav', 'compression', 'spatial_transform',
    'hetero_fusion', 'hetero_decoder', 'anchor_number'.  state_dict keys of the parts built here are the reference's
    (`fusion_net.*`, `decoder.*`, `cls_head.*`, `reg_head.*`); encoder keys exist when encoders are passed in."""

    def __init__(self, config, camera_encoder=None, lidar_encoder=None):
        super().__init__()
        self.max_cav = config.get('max_cav', 5)
        if config.get('compression', 0) > 0:
            raise NotImplementedError("compression > 0 (NaiveCompressor) is not on the shipped HM-ViT path")
        if 'hetero_decoder' not in config:
            raise NotImplementedError("only the hetero_decoder head of the shipped yaml is built")
        st = config['spatial_transform']
        self.downsample_rate = st['downsample_rate']
        self.discrete_ratio = st['voxel_size'][0]
        self.use_roi_mask = st['use_roi_mask']
        self.camera_encoder = camera_encoder
        self.lidar_encoder = lidar_encoder
        self.fusion_net = HeteroFusion(config['hetero_fusion'])
        self.use_hetero_decoder = True
        self.decoder = HeteroDecoder(config['hetero_decoder'])
        # registered by the reference even with the hetero decoder (bevformer_point_pillar_hetero.py:72-75); unused then
        self.cls_head = nn.Conv2d(256, config['anchor_number'], kernel_size=1)
        self.reg_head = nn.Conv2d(256, 7 * config['anchor_number'], kernel_size=1)

    def forward_features(self, camera_features, lidar_features, mode, record_len, pairwise_t_matrix):
        """The path behind the encoders (bevformer_point_pillar_hetero.py:113-134).  camera_features (n_cam, 256, H, W) /
        lidar_features (n_lidar, 256, H, W) in agent order within their modality (either may be None), mode (B, L),
        record_len (B,), pairwise_t_matrix (B, L, L, 4, 4)."""
        mode = mode.to(torch.int)
        max_cav = mode.shape[1]
        mode_unpack = unpad_mode_encoding(mode, record_len)
        x = combine_features(camera_features, lidar_features, mode_unpack, record_len)
        x, mask = regroup(x, record_len, max_cav)
        x = self.fusion_net(x, pairwise_t_matrix, mode, record_len, mask)
        psm, rm = self.decoder(x.unsqueeze(1), mode, use_upsample=False)
        return {'psm': psm, 'rm': rm}

    def forward(self, batch):
        mode = batch['mode'].to(torch.int)
        record_len = batch['record_len']
        mode_unpack = unpad_mode_encoding(mode, record_len)
        camera_features = lidar_features = None
        if not bool(torch.all(mode_unpack == 1)):
            if self.camera_encoder is None:
                raise NotImplementedError("no camera encoder was passed in (see hm-vit_b200/encoders.py::CvtCameraEncoder); "
                                          "call forward_features with precomputed BEV features")
            camera_features = self.camera_encoder(self.extract_camera_input(batch, mode_unpack))
        if not bool(torch.all(mode_unpack == 0)):
            if self.lidar_encoder is None:
                raise NotImplementedError("no lidar encoder was passed in (see hm-vit_b200/encoders.py::PointPillar); "
                                          "call forward_features with precomputed BEV features")
            lidar_features = self.lidar_encoder(self.extract_lidar_input(batch, mode_unpack))
        return self.forward_features(camera_features, lidar_features, mode, record_len, batch['pairwise_t_matrix'])

    # ---- encoder input selection (base_camera_lidar_intermediate.py:19-66) ----
    @staticmethod
    def extract_camera_input(batch, mode_unpack):
        sel = mode_unpack == 0
        return {k: batch[k][sel, ...] for k in ('camera', 'intrinsic', 'extrinsic', 'cav2cam_extrinsic')}

    @staticmethod
    def extract_lidar_input(batch, mode_unpack):
        """Voxels of the LiDAR agents, batch index renumbered to the agent's rank among the LiDAR agents.  Unlike the
        reference (:49) the caller's voxel_coords tensor is not modified in place."""
        pl = batch['processed_lidar']
        coords = pl['voxel_coords']
        lidar_rank = torch.cumsum((mode_unpack == 1).to(torch.long), 0) - 1           # rank of agent i among the LiDAR agents
        agent = coords[:, 0].long()
        keep = mode_unpack[agent] == 1
        # stable order by new batch index, like the reference's per-agent concatenation
        order = torch.sort(lidar_rank[agent[keep]], stable=True).indices
        new_coords = coords[keep].clone()
        new_coords[:, 0] = lidar_rank[agent[keep]].to(coords.dtype)
        return {'processed_lidar': {'voxel_features': pl['voxel_features'][keep][order],
                                    'voxel_coords': new_coords[order],
                                    'voxel_num_points': pl['voxel_num_points'][keep][order]}}
