"""BEV encoders in front of the fusion hot path ("library-backed first"): drop-in mirrors of
    PointPillar           /root/reference/opencood/models/point_pillar.py:9-61
      PillarVFE / PFNLayer  /root/reference/opencood/models/sub_modules/pillar_vfe.py:10-146
      PointPillarScatter    /root/reference/opencood/models/sub_modules/point_pillar_scatter.py:5-48
      BaseBEVBackbone       /root/reference/opencood/models/backbones/base_bev_backbone.py:6-122
      DownsampleConv        /root/reference/opencood/models/sub_modules/downsample_conv.py:9-51
    ResnetEncoder         /root/reference/opencood/models/backbones/resnet_ms.py:8-92
    CrossViewModule       /root/reference/opencood/models/sub_modules/cvt_modules.py:283-330
      CrossViewAttention :166-280, CrossAttention :96-163, BEVEmbedding :42-93
with the reference's constructor configs and state_dict keys (a reference checkpoint loads with strict=True; the golden
generator `tests/golden/make_golden_encoders.py` loads the SAME synthetic state dict into the unmodified reference modules).

These modules are OUTSIDE the hot path `north_star` names: the arithmetic is torch / cuDNN library code on whatever device the
tensors live on (convolutions, matmuls), not hand-written kernels (one exception: PillarVFE + PointPillarScatter run as ONE
hand-written kernel, `hmvit_pillar_scatter` in csrc/pillar.cuh, in eval mode on CUDA) -- SURVEY 8 f-3 ranks them "standard conv / attention,
library-backed first".  What is ours is the host logic around the library calls: the pillar features are built and scattered
without the reference's per-agent Python loops, the camera attention is evaluated in bounded chunks of agents instead of one
(b, heads, Q, n K) tensor for the whole batch, and `CvtCameraEncoder` / `build_config3_model` compose BASELINE config 3 (CVT camera
branch + PointPillar + fusion + detection decoder) behind `BevformerPointPillarHetero`, which the reference does not ship as
one model (its HM-ViT camera branch is BEVFormer on mmcv; `CrossViewTransformer.forward` calls an undefined `seg_head`,
cross_view_transformer.py:48).
"""
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# PointPillar
# ------------------------------------------------------------------------------------------------
class PFNLayer(nn.Module):
    """Linear (+ BatchNorm1d over the channel) + ReLU + max over the points of a pillar (pillar_vfe.py:10-54)."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False):
        super().__init__()
        self.last_vfe, self.use_norm = last_layer, use_norm
        width = out_channels if last_layer else out_channels // 2
        self.linear = nn.Linear(in_channels, width, bias=not use_norm)
        if use_norm:
            self.norm = nn.BatchNorm1d(width, eps=1e-3, momentum=0.01)

    def forward(self, pts):                                   # (M, P, in)
        y = self.linear(pts)
        if self.use_norm:
            y = self.norm(y.transpose(1, 2)).transpose(1, 2)
        y = torch.relu(y)
        top = y.amax(dim=1, keepdim=True)
        if self.last_vfe:
            return top
        return torch.cat([y, top.expand(-1, y.shape[1], -1)], dim=2)


class PillarVFE(nn.Module):
    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range):
        super().__init__()
        self.use_norm = model_cfg['use_norm']
        self.with_distance = model_cfg['with_distance']
        self.use_absolute_xyz = model_cfg['use_absolute_xyz']
        self.num_filters = list(model_cfg['num_filters'])
        assert len(self.num_filters) > 0
        width_in = num_point_features + (6 if self.use_absolute_xyz else 3) + (1 if self.with_distance else 0)
        widths = [width_in] + self.num_filters
        self.pfn_layers = nn.ModuleList(PFNLayer(widths[i], widths[i + 1], self.use_norm, last_layer=i == len(widths) - 2)
                                        for i in range(len(widths) - 1))
        # centre of voxel (x, y, z) = index * size + (size / 2 + range_min)
        self.voxel_xyz = [float(v) for v in voxel_size[:3]]
        self.offset_xyz = [self.voxel_xyz[k] / 2 + float(point_cloud_range[k]) for k in range(3)]

    def get_output_feature_dim(self):
        return self.num_filters[-1]

    def forward(self, batch_dict):
        pts, npts, coords = batch_dict['voxel_features'], batch_dict['voxel_num_points'], batch_dict['voxel_coords']
        M, P, _ = pts.shape
        xyz = pts[..., :3]
        mean = xyz.sum(dim=1, keepdim=True) / npts.to(pts.dtype).view(M, 1, 1)
        # coords columns are (agent, z, y, x)
        centre = torch.stack([coords[:, 3 - k].to(pts.dtype) * self.voxel_xyz[k] + self.offset_xyz[k] for k in range(3)], dim=1)
        parts = [pts if self.use_absolute_xyz else pts[..., 3:], xyz - mean, xyz - centre[:, None, :]]
        if self.with_distance:
            parts.append(xyz.norm(dim=2, keepdim=True))
        live = torch.arange(P, device=pts.device)[None, :] < npts.view(M, 1)
        feats = torch.cat(parts, dim=-1) * live[..., None].to(pts.dtype)      # padded point slots contribute zeros
        for pfn in self.pfn_layers:
            feats = pfn(feats)
        batch_dict['pillar_features'] = feats.squeeze(1)
        return batch_dict


class PointPillarScatter(nn.Module):
    """Pillars -> dense (n, C, ny, nx) canvas (point_pillar_scatter.py:15-48) as ONE indexed store.  The number of agents is
    `batch_dict['batch_size']` when the caller knows it (no device -> host read), else max(agent index) + 1 like the reference."""

    def __init__(self, model_cfg):
        super().__init__()
        self.num_bev_features = model_cfg['num_features']
        self.nx, self.ny, self.nz = model_cfg['grid_size']
        assert self.nz == 1

    def forward(self, batch_dict):
        pillars, coords = batch_dict['pillar_features'], batch_dict['voxel_coords']
        n = batch_dict.get('batch_size')
        if n is None:
            n = int(coords[:, 0].max()) + 1
        cells = self.ny * self.nx
        cell = coords[:, 1] + coords[:, 2] * self.nx + coords[:, 3]
        flat = torch.zeros(n * cells, self.num_bev_features, dtype=pillars.dtype, device=pillars.device)
        flat[(coords[:, 0] * cells + cell).long()] = pillars
        batch_dict['spatial_features'] = flat.view(n, self.ny, self.nx, self.num_bev_features).permute(0, 3, 1, 2).contiguous()
        return batch_dict


def _conv_bn_relu(c_in, c_out, **kw):
    return [nn.Conv2d(c_in, c_out, bias=False, **kw), nn.BatchNorm2d(c_out, eps=1e-3, momentum=0.01), nn.ReLU()]


class BaseBEVBackbone(nn.Module):
    """Strided conv levels + per-level up-sampling to a common resolution + channel concat (base_bev_backbone.py:6-122)."""

    def __init__(self, model_cfg, input_channels):
        super().__init__()
        nums = list(model_cfg.get('layer_nums', []))
        strides = list(model_cfg.get('layer_strides', []))
        widths = list(model_cfg.get('num_filters', []))
        assert len(nums) == len(strides) == len(widths)
        ups = list(model_cfg.get('upsample_strides', []))
        up_widths = list(model_cfg.get('num_upsample_filter', []))
        assert len(ups) == len(up_widths)
        self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList()
        c_prev = input_channels
        for lvl, (n_extra, stride, c) in enumerate(zip(nums, strides, widths)):
            layers = [nn.ZeroPad2d(1)] + _conv_bn_relu(c_prev, c, kernel_size=3, stride=stride, padding=0)
            for _ in range(n_extra):
                layers += _conv_bn_relu(c, c, kernel_size=3, padding=1)
            self.blocks.append(nn.Sequential(*layers))
            c_prev = c
            if ups:
                u = ups[lvl]
                if u >= 1:
                    self.deblocks.append(nn.Sequential(nn.ConvTranspose2d(c, up_widths[lvl], u, stride=u, bias=False),
                                                       nn.BatchNorm2d(up_widths[lvl], eps=1e-3, momentum=0.01), nn.ReLU()))
                else:
                    k = int(round(1 / u))
                    self.deblocks.append(nn.Sequential(*_conv_bn_relu(c, up_widths[lvl], kernel_size=k, stride=k)))
        c_cat = sum(up_widths)
        if len(ups) > len(nums):
            self.deblocks.append(nn.Sequential(nn.ConvTranspose2d(c_cat, c_cat, ups[-1], stride=ups[-1], bias=False),
                                               nn.BatchNorm2d(c_cat, eps=1e-3, momentum=0.01), nn.ReLU()))
        self.num_bev_features = c_cat

    def forward(self, data_dict):
        x = data_dict['spatial_features']
        full = x.shape[2]
        levels = []
        for lvl, block in enumerate(self.blocks):
            x = block(x)
            data_dict['spatial_features_%dx' % int(full / x.shape[2])] = x
            levels.append(self.deblocks[lvl](x) if len(self.deblocks) > 0 else x)
        if levels:
            x = levels[0] if len(levels) == 1 else torch.cat(levels, dim=1)
        if len(self.deblocks) > len(self.blocks):
            x = self.deblocks[-1](x)
        data_dict['spatial_features_2d'] = x
        return data_dict


class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding):
        super().__init__()
        self.double_conv = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding),
                                         nn.ReLU(inplace=True),
                                         nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
                                         nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class DownsampleConv(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList()
        c = config['input_dim']
        for k, dim, stride, pad in zip(config['kernal_size'], config['dim'], config['stride'], config['padding']):
            self.layers.append(DoubleConv(c, dim, kernel_size=k, stride=stride, padding=pad))
            c = dim

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class PointPillar(nn.Module):
    """`args` = the `lidar` block of the yaml (hypes_yaml/opcl/bevformer_point_pillar_hetero.yaml:124-150) with
    `point_pillar_scatter.grid_size` filled in by the caller like the reference's dataset code does."""

    def __init__(self, args):
        super().__init__()
        self.pillar_vfe = PillarVFE(args['pillar_vfe'], num_point_features=4, voxel_size=args['voxel_size'],
                                    point_cloud_range=args['lidar_range'])
        self.scatter = PointPillarScatter(args['point_pillar_scatter'])
        self.backbone = BaseBEVBackbone(args['base_bev_backbone'], 64)
        self.shrink_flag = 'shrink_header' in args
        if self.shrink_flag:
            self.shrink_conv = DownsampleConv(args['shrink_header'])
        self.cls_head = nn.Conv2d(args['cls_head_dim'], args['anchor_number'], kernel_size=1)
        self.reg_head = nn.Conv2d(args['cls_head_dim'], 7 * args['anchor_number'], kernel_size=1)
        self.return_features = False
        self.fused_front_end = True            # CUDA + eval: PillarVFE + scatter in one hand-written kernel
        # canvas in torch's channels_last format: removes cuDNN's NCHW <-> NHWC passes around its tensor-core convolutions, but
        # its NHWC BatchNorm inference kernel is 3.6 x slower than the NCHW one (measured: 3.63 vs 3.13 ms for the branch)
        self.channels_last = False

    def set_return_features(self):
        self.return_features = True
        return self

    def _fused_front_end(self, d):
        """PillarVFE + PointPillarScatter as ONE hand-written kernel (`hmvit_pillar_scatter`, csrc/pillar.cuh) when the
        configuration is the shipped yaml's (one PFN layer with BatchNorm, absolute xyz, no distance feature, <= 32 point
        slots, 64 features), the module is in eval mode and the voxels are on the GPU; None otherwise (torch path)."""
        vfe, sc = self.pillar_vfe, self.scatter
        pts = d['voxel_features']
        if (self.training or not pts.is_cuda or pts.dtype != torch.float32 or len(vfe.pfn_layers) != 1 or not vfe.use_norm
                or not vfe.use_absolute_xyz or vfe.with_distance or pts.shape[1] > 32 or pts.shape[2] != 4
                or vfe.num_filters[-1] != 64 or sc.num_bev_features != 64 or torch.is_grad_enabled()):
            return None
        from . import ops
        pfn = vfe.pfn_layers[0]
        s = pfn.norm.weight.float() / torch.sqrt(pfn.norm.running_var.float() + pfn.norm.eps)
        w = (pfn.linear.weight.float() * s[:, None]).contiguous()                         # BatchNorm(eval) folded: [64][10]
        b = (pfn.norm.bias.float() - pfn.norm.running_mean.float() * s).contiguous()
        n = d.get('batch_size')
        if n is None:
            n = int(d['voxel_coords'][:, 0].max()) + 1
        return ops.pillar_scatter(voxel_features=pts.contiguous(), voxel_coords=d['voxel_coords'].to(torch.int32).contiguous(),
                                  voxel_num_points=d['voxel_num_points'].to(torch.int32).contiguous(), w=w, b=b,
                                  voxel_size=vfe.voxel_xyz, offset=vfe.offset_xyz, nx=sc.nx, ny=sc.ny, n_agents=n,
                                  channels_last=self.channels_last)

    def forward(self, data_dict):
        pl = data_dict['processed_lidar']
        d = {'voxel_features': pl['voxel_features'], 'voxel_coords': pl['voxel_coords'], 'voxel_num_points': pl['voxel_num_points']}
        if 'batch_size' in data_dict:
            d['batch_size'] = data_dict['batch_size']
        canvas = self._fused_front_end(d) if self.fused_front_end else None
        if canvas is not None:
            d['spatial_features'] = canvas
        else:
            d = self.scatter(self.pillar_vfe(d))
        feat = self.backbone(d)['spatial_features_2d']
        if self.shrink_flag:
            feat = self.shrink_conv(feat)
        if self.return_features:
            return feat
        return {'psm': self.cls_head(feat), 'rm': self.reg_head(feat)}


# ------------------------------------------------------------------------------------------------
# CVT camera branch
# ------------------------------------------------------------------------------------------------
def _conv_out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


class ResnetEncoder(nn.Module):
    """torchvision ResNet trunk, picks of the four stage outputs (resnet_ms.py:8-92).  (B, L, M, H, W, 3) images ->
    (B, L, M, C, h, w) per picked stage.  `output_shapes` is computed from the layer arithmetic instead of a dummy forward."""

    _STAGE_WIDTH = {18: 64, 34: 64, 50: 256, 101: 256, 152: 256}

    def __init__(self, params):
        super().__init__()
        import torchvision.models as tvm
        self.num_layers = params['num_layers']
        self.idx_pick = params['id_pick']
        if self.num_layers not in self._STAGE_WIDTH:
            raise ValueError("{} is not a valid number of resnet layers".format(self.num_layers))
        weights = 'DEFAULT' if params.get('pretrained', False) else None       # needs the torchvision weight cache; no network here
        self.encoder = getattr(tvm, 'resnet%d' % self.num_layers)(weights=weights)
        h, w = params['image_height'], params['image_width']
        h, w = _conv_out(_conv_out(h, 7, 2, 3), 3, 2, 1), _conv_out(_conv_out(w, 7, 2, 3), 3, 2, 1)      # conv1, maxpool
        shapes, c = [], self._STAGE_WIDTH[self.num_layers]
        for stage in range(4):
            if stage:
                h, w, c = _conv_out(h, 3, 2, 1), _conv_out(w, 3, 2, 1), c * 2
            shapes.append(torch.Size([1, 1, 1, c, h, w]))
        self.output_shapes = [shapes[i] for i in self.idx_pick] if isinstance(self.idx_pick, list) else shapes[self.idx_pick]

    def forward(self, input_images):
        b, l, m, h, w, c = input_images.shape
        x = input_images.reshape(b * l * m, h, w, c).permute(0, 3, 1, 2).contiguous()
        e = self.encoder
        x = e.maxpool(e.relu(e.bn1(e.conv1(x))))
        outs = []
        for stage in (e.layer1, e.layer2, e.layer3, e.layer4):
            x = stage(x)
            outs.append(x.view(b, l, m, *x.shape[1:]))
        return [outs[i] for i in self.idx_pick] if isinstance(self.idx_pick, list) else outs[self.idx_pick]


def _unit_plane(height: int, width: int) -> torch.Tensor:
    """The reference's `generate_grid` (cvt_modules.py:16-28) INCLUDING its axis convention: the result has shape
    (3, width, height) with plane[0, i, j] = j / (height - 1), plane[1, i, j] = i / (width - 1), plane[2] = 1 (its meshgrid runs
    over (xs, ys) with 'ij' indexing); callers flatten it and re-read the flat index as (h, w), which is the identity for the
    square image feature maps and a re-interpretation for the 48 x 176 BEV grid.  Reproduced so that checkpoints mean the same."""
    along_w = torch.linspace(0, 1, width)
    along_h = torch.linspace(0, 1, height)
    return torch.stack([along_h[None, :].expand(width, height), along_w[:, None].expand(width, height),
                        torch.ones(width, height)]).contiguous()


class BEVEmbedding(nn.Module):
    """Learned BEV prior + the metric coordinates of its cells (cvt_modules.py:42-93)."""

    def __init__(self, dim, sigma, bev_height, bev_width, h_meters, w_meters, offset, decoder_blocks):
        super().__init__()
        h = bev_height // (2 ** len(decoder_blocks))
        w = bev_width // (2 ** len(decoder_blocks))
        plane = _unit_plane(h, w)
        plane[0] *= bev_width
        plane[1] *= bev_height
        sh, sw = bev_height / h_meters, bev_width / w_meters
        view = torch.tensor([[0., -sw, bev_width / 2.], [-sh, 0., bev_height * offset + bev_height / 2.], [0., 0., 1.]])
        grid = (torch.linalg.inv(view) @ plane.reshape(3, -1)).reshape(3, h, w)
        self.register_buffer('grid', grid, persistent=False)
        self.learned_features = nn.Parameter(sigma * torch.randn(dim, h, w))

    def get_prior(self):
        return self.learned_features


class CrossAttention(nn.Module):
    """BEV queries x image keys of all cameras of an agent, one softmax across cameras and pixels (cvt_modules.py:96-163)."""

    def __init__(self, dim, heads, dim_head, qkv_bias, norm=nn.LayerNorm):
        super().__init__()
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        inner = heads * dim_head
        self.to_q = nn.Sequential(norm(dim), nn.Linear(dim, inner, bias=qkv_bias))
        self.to_k = nn.Sequential(norm(dim), nn.Linear(dim, inner, bias=qkv_bias))
        self.to_v = nn.Sequential(norm(dim), nn.Linear(dim, inner, bias=qkv_bias))
        self.proj = nn.Linear(inner, dim)
        self.prenorm = norm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, 2 * dim), nn.GELU(), nn.Linear(2 * dim, dim))
        self.postnorm = norm(dim)
        self.max_logits = 1 << 30              # elements of one logits tensor; agents are processed in chunks below this
        self.fused = True                      # CUDA: per-camera fused attention (library kernel) + log-sum-exp merge

    def _attend_exact(self, q, k, v):
        """fp32 reference form: logits of all cameras of an agent materialised, one softmax over (camera, pixel)."""
        b, n, Q, m, _ = q.shape
        step = max(1, self.max_logits // max(1, m * Q * n * k.shape[2]))
        outs = []
        for s in range(0, b, step):
            logits = self.scale * torch.einsum('bnqmd,bnkmd->bmqnk', q[s:s + step], k[s:s + step])
            att = logits.flatten(3).softmax(dim=-1)                                        # over (camera, pixel)
            outs.append(torch.einsum('bmqk,bkmd->bqmd', att, v[s:s + step].flatten(1, 2)).flatten(2))
        return torch.cat(outs) if len(outs) > 1 else outs[0]

    def _attend_fused(self, q, k, v):
        """The same softmax without the (Q x n K) logits in HBM: one memory-efficient attention call per (agent, camera) that
        also returns the log-sum-exp of its logits, then the cameras are merged with weights exp(lse_n - logsumexp_n lse_n).
        Library kernel (torch's cutlass FMHA: TF32 tensor-core contractions for fp32 inputs)."""
        b, n, Q, m, dh = q.shape

        def bh(t):                                                                         # (b, n, L, m, d) -> (b n, m, L, d)
            return t.permute(0, 1, 3, 2, 4).reshape(b * n, m, t.shape[2], dh).contiguous()
        out, lse = torch.ops.aten._scaled_dot_product_efficient_attention(bh(q), bh(k), bh(v), None, True, scale=self.scale)[:2]
        w = lse[..., :Q].reshape(b, n, m, Q).softmax(dim=1)                                # weight of a camera for (head, query)
        out = (out.reshape(b, n, m, Q, dh) * w[..., None]).sum(dim=1)                      # (b, m, Q, d)
        return out.transpose(1, 2).reshape(b, Q, m * dh)

    def forward(self, q, k, v, skip=None):
        """q (b, n, d, H, W), k / v (b, n, d, h, w) -> (b, d, H, W)."""
        b, n, _, H, W = q.shape
        m, dh = self.heads, self.dim_head
        q = self.to_q(q.flatten(3).transpose(2, 3)).view(b, n, H * W, m, dh)
        k = self.to_k(k.flatten(3).transpose(2, 3)).view(b, n, -1, m, dh)
        v = self.to_v(v.flatten(3).transpose(2, 3)).view(b, n, -1, m, dh)
        a = self._attend_fused(q, k, v) if (self.fused and q.is_cuda and not torch.is_grad_enabled()) else self._attend_exact(q, k, v)
        z = self.proj(a)
        if skip is not None:
            z = z + skip.flatten(2).transpose(1, 2)
        z = self.prenorm(z)
        z = self.postnorm(z + self.mlp(z))
        return z.transpose(1, 2).reshape(b, -1, H, W)


class CrossViewAttention(nn.Module):
    """Camera-aware positional embeddings for keys (image rays) and queries (BEV cells), then CrossAttention
    (cvt_modules.py:166-280)."""

    def __init__(self, feat_height, feat_width, feat_dim, dim, config):
        super().__init__()
        plane = _unit_plane(feat_height, feat_width)[None, None].clone()           # 1 1 3 fw fh
        plane[:, :, 0] *= config['image_width']
        plane[:, :, 1] *= config['image_height']
        self.register_buffer('image_plane', plane, persistent=False)

        def bn_relu_conv():
            return nn.Sequential(nn.BatchNorm2d(feat_dim), nn.ReLU(), nn.Conv2d(feat_dim, dim, 1, bias=False))
        self.feature_linear = bn_relu_conv()
        self.feature_proj = None if config['no_image_features'] else bn_relu_conv()
        self.bev_embed = nn.Conv2d(2, dim, 1)
        self.img_embed = nn.Conv2d(4, dim, 1, bias=False)
        self.cam_embed = nn.Conv2d(4, dim, 1, bias=False)
        self.cross_attend = CrossAttention(dim, config['heads'], config['dim_head'], config['qkv_bias'])
        self.skip = config['skip']

    def forward(self, x, bev, feature, I_inv, E_inv):
        """x (b, d, H, W) BEV state, feature (b, n, c, h, w), I_inv (b, n, 3, 3), E_inv (b, n, 4, 4) -> (b, d, H, W)."""
        b, n = feature.shape[:2]
        ph, pw = self.image_plane.shape[-2:]
        cam_pos = self.cam_embed(E_inv[..., -1:].reshape(b * n, 4, 1, 1))                       # (bn, d, 1, 1) camera centre
        rays = I_inv @ self.image_plane.flatten(3)                                              # (b, n, 3, hw)
        rays = E_inv @ F.pad(rays, (0, 0, 0, 1), value=1.0)                                     # (b, n, 4, hw)
        img_pos = self.img_embed(rays.reshape(b * n, 4, ph, pw)) - cam_pos
        img_pos = img_pos / (img_pos.norm(dim=1, keepdim=True) + 1e-7)
        bev_pos = self.bev_embed(bev.grid[:2][None]) - cam_pos                                  # (bn, d, H, W)
        bev_pos = bev_pos / (bev_pos.norm(dim=1, keepdim=True) + 1e-7)
        feat = feature.flatten(0, 1)
        key = img_pos if self.feature_proj is None else img_pos + self.feature_proj(feat)
        val = self.feature_linear(feat)
        query = bev_pos.view(b, n, *bev_pos.shape[1:]) + x[:, None]
        return self.cross_attend(query, key.view(b, n, *key.shape[1:]), val.view(b, n, *val.shape[1:]),
                                 skip=x if self.skip else None)


class CrossViewModule(nn.Module):
    def __init__(self, config):
        super().__init__()
        from torchvision.models.resnet import Bottleneck
        dim = config['dim']
        self.backbone_output_shape = config['backbone_output_shape']
        assert len(config['middle']) == len(self.backbone_output_shape)
        self.bev_embedding = BEVEmbedding(dim, **config['bev_embedding'])
        self.cross_views = nn.ModuleList(CrossViewAttention(s[-2], s[-1], s[-3], dim, config['cross_view'])
                                         for s in self.backbone_output_shape)
        self.layers = nn.ModuleList(nn.Sequential(*[Bottleneck(dim, dim // 4) for _ in range(k)]) for k in config['middle'])

    def forward(self, batch):
        b, l = batch['inputs'].shape[:2]
        I_inv = torch.linalg.inv(batch['intrinsic'].flatten(0, 1))
        E = batch['extrinsic'].flatten(0, 1)
        x = self.bev_embedding.get_prior()[None].expand(b * l, -1, -1, -1)
        for cross_view, feature, layer in zip(self.cross_views, batch['features'], self.layers):
            x = layer(cross_view(x, self.bev_embedding, feature.flatten(0, 1), I_inv, E))
        return x.view(b, l, *x.shape[1:])


class CvtCameraEncoder(nn.Module):
    """Camera branch of BASELINE config 3: `ResnetEncoder` + `CrossViewModule` with the attribute names of the reference's
    `CrossViewTransformer` (`encoder`, `cvm`; cross_view_transformer.py:17-24) and the input dict of
    `BaseCameraLiDARIntermediate.extract_camera_input`: camera (n, m, h, w, 3), intrinsic (n, m, 3, 3), extrinsic (n, m, 4, 4)
    -> (n, dim, H, W)."""

    def __init__(self, config):
        super().__init__()
        self.encoder = ResnetEncoder(config['encoder'])
        cvm = dict(config['cvm'])
        cvm['backbone_output_shape'] = self.encoder.output_shapes
        self.cvm = CrossViewModule(cvm)

    def forward(self, batch):
        cam = batch['camera'].unsqueeze(1)
        feats = self.encoder(cam)
        out = self.cvm({'inputs': cam, 'features': feats if isinstance(feats, list) else [feats],
                        'intrinsic': batch['intrinsic'].unsqueeze(1), 'extrinsic': batch['extrinsic'].unsqueeze(1)})
        return out[:, 0]


def config3_args(bev_h=48, bev_w=176, image=512, resnet=34, anchor_number=2, max_cav=5, voxel=0.4, z_range=(-3.0, 1.0)) -> Dict:
    """Model arguments of BASELINE config 3: the `lidar` / `hetero_*` blocks of the shipped yaml
    (hypes_yaml/opcl/bevformer_point_pillar_hetero.yaml:81-156) + the CVT blocks of hypes_yaml/opcamera/cvt.yaml:60-90 at
    dim 256 on the (bev_h, bev_w) grid.  The LiDAR grid is 4x the BEV grid at `voxel` metres (704 x 192 at 48 x 176)."""
    half_x, half_y = 2 * bev_w * voxel, 2 * bev_h * voxel
    st = {'downsample_rate': 4, 'voxel_size': [voxel, voxel, z_range[1] - z_range[0]], 'use_roi_mask': True}
    return {
        'max_cav': max_cav, 'anchor_number': anchor_number, 'compression': 0, 'spatial_transform': st,
        'hetero_fusion': {'num_iters': 2, 'spatial_transform': st,
                          'hetero_fusion_block': {'spatial_transform': st, 'architect_mode': 'sequential', 'input_dim': 256,
                                                  'mlp_dim': 256, 'agent_size': max_cav, 'window_size': 8, 'dim_head': 32,
                                                  'drop_out': 0.1, 'mask': True}},
        'hetero_decoder': {'input_dim': 256, 'num_layer': 2, 'num_ch_dec': [256, 256], 'anchor_number': anchor_number},
        'camera': {'encoder': {'num_layers': resnet, 'pretrained': False, 'image_width': image, 'image_height': image, 'id_pick': [1, 3]},
                   'cvm': {'dim': 256, 'middle': [2, 2],
                           'bev_embedding': {'sigma': 1.0, 'bev_height': bev_h, 'bev_width': bev_w, 'h_meters': 2 * half_y,
                                             'w_meters': 2 * half_x, 'offset': 0.0, 'decoder_blocks': []},
                           'cross_view': {'image_height': image, 'image_width': image, 'no_image_features': False, 'skip': True,
                                          'heads': 4, 'dim_head': 32, 'qkv_bias': True}}},
        'lidar': {'voxel_size': st['voxel_size'], 'lidar_range': [-half_x, -half_y, z_range[0], half_x, half_y, z_range[1]],
                  'anchor_number': anchor_number,
                  'pillar_vfe': {'use_norm': True, 'with_distance': False, 'use_absolute_xyz': True, 'num_filters': [64]},
                  'point_pillar_scatter': {'num_features': 64, 'grid_size': [4 * bev_w, 4 * bev_h, 1]},
                  'base_bev_backbone': {'layer_nums': [3, 5, 8], 'layer_strides': [2, 2, 2], 'num_filters': [64, 128, 256],
                                        'upsample_strides': [1, 2, 4], 'num_upsample_filter': [128, 128, 128]},
                  'shrink_header': {'kernal_size': [3], 'stride': [2], 'padding': [1], 'dim': [256], 'input_dim': 384},
                  'cls_head_dim': 256},
    }


def build_config3_model(args: Dict):
    """BASELINE config 3 as one module: CVT camera encoder + PointPillar behind `BevformerPointPillarHetero` (model.py), i.e.
    encoders (library) -> combine / regroup -> HeteroFusion (hmvit_fusion_forward) -> HeteroDecoder (hmvit_decoder_forward)."""
    from .model import BevformerPointPillarHetero
    return BevformerPointPillarHetero(args, camera_encoder=CvtCameraEncoder(args['camera']),
                                      lidar_encoder=PointPillar(args['lidar']).set_return_features())
