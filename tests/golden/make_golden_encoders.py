"""Reference outputs for the BEV encoders and BASELINE config 3 (run in the build container only; needs /root/reference):

    python tests/golden/make_golden_encoders.py      -> tests/golden/encoders.npz

The UNMODIFIED reference modules (`PointPillar`, `ResnetEncoder`, `CrossViewModule`, `HeteroFusion`, `HeteroDecoder`, `regroup`)
are constructed with the configs of `hm-vit_b200/encoders.py::config3_args` at a small grid, loaded (strict=True) with the
synthetic state dicts of `tests/enc_synth.py` / `oracle.synth_*`, and run in eval mode on the synthetic inputs of
`tests/enc_synth.py`.  Stored: the encoder features, and for the composite the per-agent BEV features, the fused feature and
the detection logits.  The tests rebuild inputs and weights from the same seeds.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import ref_import  # noqa: E402
import hmvit_loader  # noqa: E402
import enc_synth as S  # noqa: E402
from oracle import hmvit_oracle as O  # noqa: E402

BEV_H, BEV_W, IMAGE = 16, 24, 64          # BEV grid (multiples of the 8 x 8 window), camera image side
C3_MODE = [[1, 0, 0], [0, 1, 0]]          # scene 0: (lidar, camera, camera); scene 1: (camera, lidar)
C3_RECORD_LEN = [3, 2]
C3_ORDER = [('l', 0), ('c', 0), ('c', 1), ('c', 2), ('l', 1)]     # agent -> row of its modality's feature tensor


def main():
    R = ref_import.load()
    from opencood.models.point_pillar import PointPillar
    from opencood.models.sub_modules.cvt_modules import CrossViewModule
    from opencood.models.backbones.resnet_ms import ResnetEncoder
    pkg = hmvit_loader.load()
    args = pkg.encoders.config3_args(bev_h=BEV_H, bev_w=BEV_W, image=IMAGE)
    out = {}
    with torch.no_grad():
        # ---- PointPillar (2 LiDAR agents) ----
        la = args['lidar']
        pp = PointPillar(la).eval()
        pp.load_state_dict(S.synth_module_state_dict(pp, 1), strict=True)
        nx, ny, _ = la['point_pillar_scatter']['grid_size']
        vox = S.synth_voxels(2, nx, ny, 400, la['lidar_range'], la['voxel_size'], seed=1)
        heads = pp({'processed_lidar': vox})
        pp.set_return_features()
        lidar_feat = pp({'processed_lidar': vox})
        out['pp_features'], out['pp_psm'], out['pp_rm'] = lidar_feat.numpy(), heads['psm'].numpy(), heads['rm'].numpy()
        print("PointPillar", tuple(lidar_feat.shape), float(lidar_feat.abs().mean()))

        # ---- CVT camera branch (3 camera agents x 2 cameras) ----
        ca = args['camera']
        enc = ResnetEncoder(dict(ca['encoder'])).eval()
        enc.load_state_dict(S.synth_module_state_dict(enc, 2), strict=True)
        cvm_cfg = dict(ca['cvm'])
        cvm_cfg['backbone_output_shape'] = enc.output_shapes
        cvm = CrossViewModule(cvm_cfg).eval()
        cvm.load_state_dict(S.synth_module_state_dict(cvm, 3), strict=True)
        cams = S.synth_cameras(3, 2, IMAGE, seed=2)
        inputs = cams['camera'].unsqueeze(1)
        feats = enc(inputs)
        cam_feat = cvm({'inputs': inputs, 'features': feats, 'intrinsic': cams['intrinsic'].unsqueeze(1),
                        'extrinsic': cams['extrinsic'].unsqueeze(1)})[:, 0]
        out['resnet_shapes'] = np.array([list(s) for s in enc.output_shapes])
        out['resnet_f0'], out['resnet_f1'] = feats[0].numpy(), feats[1].numpy()
        out['cvt_features'] = cam_feat.numpy()
        print("CVT", tuple(cam_feat.shape), float(cam_feat.abs().mean()), [tuple(f.shape) for f in feats])

        # ---- BASELINE config 3 composite: 2 scenes, agents (lidar, camera, camera) and (camera, lidar) ----
        mode = torch.tensor(C3_MODE)
        record_len = torch.tensor(C3_RECORD_LEN)
        _, T, _, _, _ = O.synth_inputs(2, 3, 256, BEV_H, BEV_W, C3_RECORD_LEN, seed=31, tx=6.0, ty=4.0)
        x = torch.stack([lidar_feat[i] if k == 'l' else cam_feat[i] for k, i in C3_ORDER])
        xg, mask = R.regroup(x, record_len, 3)
        cfg = O.default_config()
        fusion = R.HeteroFusion(cfg).eval()
        fusion.load_state_dict(O.synth_state_dict(cfg, 0), strict=True)
        fused = fusion(xg, T, mode, record_len, mask)
        dec = R.HeteroDecoder(args['hetero_decoder']).eval()
        dec.load_state_dict(O.synth_decoder_state_dict(1), strict=True)
        psm, rm = dec(fused.unsqueeze(1), mode, use_upsample=False)
        out['c3_T'], out['c3_fused'], out['c3_psm'], out['c3_rm'] = T.numpy(), fused.numpy(), psm.numpy(), rm.numpy()
        print("config 3", tuple(fused.shape), float(fused.abs().mean()), tuple(psm.shape), float(psm.abs().mean()))
    np.savez_compressed(os.path.join(HERE, "encoders.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print("wrote encoders.npz", os.path.getsize(os.path.join(HERE, "encoders.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
