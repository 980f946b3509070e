"""Deterministic synthetic weights and inputs for the BEV encoders.  Shared by the golden generator
(`tests/golden/make_golden_encoders.py`, which loads them into the UNMODIFIED reference modules with strict=True) and the
tests (which load them into `hm-vit_b200/encoders.py`): a tensor's values depend only on (seed, key name, shape)."""
import zlib

import torch


def synth_module_state_dict(module, seed=0):
    """Every floating entry of `module.state_dict()` drawn from a generator seeded by crc32(key): conv / linear weights
    N(0, 1/fan_in) (He-like 1.4x for convolutions so activations do not fade through deep ReLU stacks), BatchNorm / LayerNorm
    scale 1 + 0.1 n, biases and running means 0.1 n, running variances 1 + 0.2 u, learned BEV prior N(0, 1)."""
    out = {}
    for key, ref in module.state_dict().items():
        if not ref.is_floating_point():
            out[key] = ref.clone()
            continue
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
        leaf = key.rsplit('.', 1)[-1]
        if leaf == 'running_var':
            t = 1.0 + 0.2 * torch.rand(ref.shape, generator=g)
        elif leaf in ('running_mean', 'bias'):
            t = 0.1 * torch.randn(ref.shape, generator=g)
        elif leaf == 'learned_features':
            t = torch.randn(ref.shape, generator=g)
        elif ref.dim() >= 2:
            fan_in = ref[0].numel()
            gain = 1.4 if ref.dim() == 4 and ref.shape[-1] > 1 else 1.0
            t = gain * torch.randn(ref.shape, generator=g) / fan_in ** 0.5
        else:                                                   # 1-D weight: a norm scale
            t = 1.0 + 0.1 * torch.randn(ref.shape, generator=g)
        out[key] = t.to(ref.dtype)
    return out


def synth_voxels(n_agents, nx, ny, per_agent, lidar_range, voxel_size, points=32, seed=0):
    """Synthetic PointPillar input like the reference's voxel generator hands it over: `per_agent` distinct non-empty pillars
    per agent, 1..points points each inside the pillar's cell (x, y, z, intensity), zero-padded; coords (agent, z, y, x)."""
    g = torch.Generator().manual_seed(4242 + seed)
    feats, coords, nums = [], [], []
    for a in range(n_agents):
        cell = torch.randperm(nx * ny, generator=g)[:per_agent]
        cy, cx = cell // nx, cell % nx
        npt = torch.randint(1, points + 1, (per_agent,), generator=g)
        u = torch.rand(per_agent, points, 3, generator=g)
        x = lidar_range[0] + (cx[:, None] + u[..., 0]) * voxel_size[0]
        y = lidar_range[1] + (cy[:, None] + u[..., 1]) * voxel_size[1]
        z = lidar_range[2] + u[..., 2] * voxel_size[2]
        inten = torch.rand(per_agent, points, generator=g)
        p = torch.stack([x, y, z, inten], dim=-1)
        p = p * (torch.arange(points)[None, :] < npt[:, None])[..., None]
        feats.append(p.float())
        coords.append(torch.stack([torch.full_like(cx, a), torch.zeros_like(cx), cy, cx], dim=1))
        nums.append(npt)
    return {'voxel_features': torch.cat(feats), 'voxel_coords': torch.cat(coords).int(), 'voxel_num_points': torch.cat(nums).int()}


def synth_cameras(n_agents, n_cam, image, seed=0):
    """Images (n, m, image, image, 3) in [0, 1], pin-hole intrinsics and camera-to-ego style extrinsics (rotation about z +
    translation) per camera."""
    g = torch.Generator().manual_seed(777 + seed)
    cam = torch.rand(n_agents, n_cam, image, image, 3, generator=g)
    K = torch.zeros(n_agents, n_cam, 3, 3)
    f = image * (0.8 + 0.4 * torch.rand(n_agents, n_cam, generator=g))
    K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2], K[..., 2, 2] = f, f, image / 2, image / 2, 1.0
    yaw = 6.2831853 * torch.rand(n_agents, n_cam, generator=g)
    E = torch.zeros(n_agents, n_cam, 4, 4)
    E[..., 0, 0], E[..., 0, 1], E[..., 1, 0], E[..., 1, 1] = yaw.cos(), -yaw.sin(), yaw.sin(), yaw.cos()
    E[..., 2, 2] = E[..., 3, 3] = 1.0
    E[..., :3, 3] = 2.0 * torch.randn(n_agents, n_cam, 3, generator=g)
    return {'camera': cam, 'intrinsic': K, 'extrinsic': E}


def config3_batch(vox, cams, order, mode, record_len, T):
    """The batch dict as the collate function lays it out for `BevformerPointPillarHetero.forward`: per-agent tensors in scene
    order (camera entries of LiDAR agents are dummies), voxel batch index = agent index over ALL agents.  `order` = per agent
    ('l' | 'c', row in its modality's synthetic input)."""
    n_agents = len(order)
    is_lidar = torch.tensor([k == 'l' for k, _ in order])
    lidar_agent = torch.nonzero(is_lidar)[:, 0]                     # agent index of the i-th LiDAR agent
    cam_agent = torch.nonzero(~is_lidar)[:, 0]
    coords = vox['voxel_coords'].clone()
    coords[:, 0] = lidar_agent[coords[:, 0].long()].to(coords.dtype)
    batch = {'mode': mode.double(), 'record_len': record_len, 'pairwise_t_matrix': T,
             'processed_lidar': {'voxel_features': vox['voxel_features'], 'voxel_coords': coords,
                                 'voxel_num_points': vox['voxel_num_points']}}
    for key, val in cams.items():
        full = torch.zeros((n_agents,) + tuple(val.shape[1:]), dtype=val.dtype)
        full[cam_agent] = val
        if key != 'camera':
            full[lidar_agent] = torch.eye(val.shape[-1])
        batch[key] = full
    batch['cav2cam_extrinsic'] = batch['extrinsic']
    return batch
