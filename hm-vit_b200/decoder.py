"""Detection decoder of the HM-ViT model on B200: drop-in mirrors of
    HeteroDecoder   /root/reference/opencood/models/sub_modules/hetero_decoder.py:7-91
    NaiveDecoder    /root/reference/opencood/models/sub_modules/naive_decoder.py:8-92
with the same constructor parameters, forward signatures and state_dict keys (`camera_decoder.decoder.<i>.*`,
`lidar_decoder.decoder.<i>.*`, `{camera,lidar}_{cls,reg}_head.*`), so a `net_epoch*.pth` of the reference loads.

All arithmetic runs in one C call (`hmvit_decoder_forward`, csrc/decoder.cuh): per scene, with the weights of the ego's
modality, 2 * num_layer x (conv3x3 + BatchNorm + ReLU) as TMA-shifted implicit GEMMs on tcgen05 and the two 1x1 heads.
Eval-mode BatchNorm is folded into the convolution here:
    BN(conv(x)) = conv(x) * s + (b - mean) * s + beta,   s = gamma / sqrt(var + eps).
There is no CPU / eager fallback.  Train mode (batch statistics) and `use_upsample=True` are outside the path the HM-ViT
model calls (`bevformer_point_pillar_hetero.py:125-126` passes use_upsample=False; SURVEY.md 8 a13 covers the fusion
backward only) and raise.
"""
from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import ops

_KERNEL_CH = 256


class NaiveDecoder(nn.Module):
    """Parameter container with the reference's registration order (naive_decoder.py:27-54)."""

    def __init__(self, params):
        super().__init__()
        self.num_ch_dec = list(params['num_ch_dec'])
        self.num_layer = params['num_layer']
        self.input_dim = params['input_dim']
        assert len(self.num_ch_dec) == self.num_layer
        convs = OrderedDict()
        for i in range(self.num_layer - 1, -1, -1):
            num_ch_in = self.input_dim if i == self.num_layer - 1 else self.num_ch_dec[i + 1]
            num_ch_out = self.num_ch_dec[i]
            convs[("upconv", i, 0)] = nn.Conv2d(num_ch_in, num_ch_out, 3, 1, 1)
            convs[("norm", i, 0)] = nn.BatchNorm2d(num_ch_out)
            convs[("relu", i, 0)] = nn.ReLU(True)
            convs[("upconv", i, 1)] = nn.Conv2d(num_ch_out, num_ch_out, 3, 1, 1)
            convs[("norm", i, 1)] = nn.BatchNorm2d(num_ch_out)
            convs[("relu", i, 1)] = nn.ReLU(True)
        self.decoder = nn.ModuleList(list(convs.values()))
        if self.input_dim != _KERNEL_CH or any(c != _KERNEL_CH for c in self.num_ch_dec):
            raise ValueError("hmvit_b200 decoder kernels are specialised for input_dim = num_ch_dec = 256 "
                             f"(got input_dim={self.input_dim}, num_ch_dec={self.num_ch_dec})")

    def folded(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """BatchNorm-folded weights of the 2 * num_layer convolutions: W [n][9][256 out][256 in] fp32 (tap = ky * 3 + kx),
        b [n][256] fp32."""
        ws, bs = [], []
        for i in range(0, len(self.decoder), 3):
            conv, bn = self.decoder[i], self.decoder[i + 1]
            s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            w = conv.weight.detach().float() * s[:, None, None, None]                     # [out][in][3][3]
            b = (conv.bias.detach().float() - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
            ws.append(w.permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1]))             # [ky*3+kx][out][in]
            bs.append(b)
        return torch.stack(ws), torch.stack(bs)

    def forward(self, x, use_upsample=True):
        raise NotImplementedError("NaiveDecoder alone is not on the HM-ViT path; use HeteroDecoder (hetero_decoder.py:42-74)")


class HeteroDecoder(nn.Module):
    def __init__(self, params):
        super().__init__()
        input_dim = params['num_ch_dec'][0]
        self.anchor_number = params['anchor_number']
        self.camera_decoder = NaiveDecoder(params)
        self.lidar_decoder = NaiveDecoder(params)
        self.camera_cls_head = nn.Conv2d(input_dim, params['anchor_number'], kernel_size=1)
        self.camera_reg_head = nn.Conv2d(input_dim, 7 * params['anchor_number'], kernel_size=1)
        self.lidar_cls_head = nn.Conv2d(input_dim, params['anchor_number'], kernel_size=1)
        self.lidar_reg_head = nn.Conv2d(input_dim, 7 * params['anchor_number'], kernel_size=1)
        if not 1 <= self.anchor_number <= 4:
            raise ValueError("hmvit_b200 decoder kernels support anchor_number in 1..4")
        self._pack_cache = None

    # ---- folded / cast weights, cached like HeteroFusionBlock.packed() ----
    def packed(self) -> Dict[str, torch.Tensor]:
        key = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._pack_cache is None or self._pack_cache[0] != key:
            wc, bc = self.camera_decoder.folded()
            wl, bl = self.lidar_decoder.folded()
            conv_w = torch.stack([wc, wl], dim=1).to(torch.float16).contiguous()          # [n][2][9][256][256]
            conv_b = torch.stack([bc, bl], dim=1).contiguous()                              # [n][2][256]
            A = self.anchor_number

            def head(cls, reg):
                return (torch.cat([cls.weight.detach().float().view(A, -1), reg.weight.detach().float().view(7 * A, -1)]),
                        torch.cat([cls.bias.detach().float(), reg.bias.detach().float()]))
            hwc, hbc = head(self.camera_cls_head, self.camera_reg_head)
            hwl, hbl = head(self.lidar_cls_head, self.lidar_reg_head)
            self._pack_cache = (key, {"conv_w": conv_w, "conv_b": conv_b,
                                      "head_w": torch.stack([hwc, hwl]).contiguous(), "head_b": torch.stack([hbc, hbl]).contiguous()})
        return self._pack_cache[1]

    def invalidate_packed(self):
        self._pack_cache = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._pack_cache = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._pack_cache = None
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode=True):
        self._pack_cache = None
        return super().train(mode)

    def forward(self, x, mode, use_upsample=True):
        """x (B, 1, C, H, W) or (B, C, H, W): the fused ego feature; mode (B, L): modality flags, column 0 = the ego's
        (hetero_decoder.py:54).  Returns psm (B, A, H, W), rm (B, 7A, H, W)."""
        if use_upsample:
            raise NotImplementedError("use_upsample=True is not on the HM-ViT path (bevformer_point_pillar_hetero.py:126)")
        if self.training:
            raise NotImplementedError("the decoder kernels implement eval-mode BatchNorm (running statistics) only")
        if x.dim() == 5:
            if x.shape[1] != 1:
                raise ValueError(f"expected the ego feature (B, 1, C, H, W), got {tuple(x.shape)}")
            x = x[:, 0]
        if x.dim() != 4 or x.shape[1] != _KERNEL_CH:
            raise ValueError(f"expected (B, {_KERNEL_CH}, H, W), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise ValueError("hmvit_b200 has no CPU path: the input must be a CUDA tensor")
        B, _, H, W = x.shape
        if H % 8 or W % 8:
            raise ValueError("H and W must be divisible by 8")
        ego_mode = mode[:, 0].to(torch.int32).contiguous()
        if bool(((ego_mode != 0) & (ego_mode != 1)).any()):
            raise ValueError("Mode but be either 1 or 0")                                   # (hetero_decoder.py:86-88)
        pk = self.packed()
        x = x.float().contiguous()
        A = self.anchor_number
        psm = torch.empty(B, A, H, W, dtype=torch.float32, device=x.device)
        rm = torch.empty(B, 7 * A, H, W, dtype=torch.float32, device=x.device)
        ws = torch.empty(ops.decoder_workspace_bytes(B, H, W) + 1024, dtype=torch.uint8, device=x.device)
        off = (-ws.data_ptr()) % 1024
        ops.decoder_forward(x=x, ego_mode=ego_mode, conv_w=pk["conv_w"], conv_b=pk["conv_b"], head_w=pk["head_w"],
                            head_b=pk["head_b"], anchor_number=A, psm=psm, rm=rm, workspace=ws[off:off + ws.numel() - 1024])
        return psm, rm
