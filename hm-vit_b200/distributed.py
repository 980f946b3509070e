"""Data-parallel training support for the fusion module.

The reference's only parallelism is DistributedDataParallel over NCCL with find_unused_parameters=True
(opencood/tools/train_camera.py:126-131): replicas + a bucketed gradient all-reduce.  The fusion forward and
backward need no communication (scenes are independent), and this module's parameter gradients only
materialise at the end of its hand-written backward (they are pulled back from the folded-weight gradients in one
step), so there is nothing to overlap inside the module: the 2.5 M fusion parameters (10 MB fp32) travel as ONE
flat bucket in ONE all-reduce over NCCL / NVLink.  A per-parameter "used" flag rides in the same bucket so that
parameters unused on every rank (aggregate_fc, weights of a modality absent from the whole global batch) keep
grad = None, as under DDP's find_unused_parameters.
"""
from __future__ import annotations

from typing import List

import torch


class FlatGradAllReduce:
    """bucket = [grad_0 | grad_1 | ... | used flags]; one all-reduce per step."""

    def __init__(self, module: torch.nn.Module, device=None):
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.total = sum(self.sizes)
        dev = device if device is not None else (self.params[0].device if self.params else torch.device("cpu"))
        self.flat = torch.zeros(self.total + len(self.params), dtype=torch.float32, device=dev)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def pack(self):
        off = 0
        flags = self.flat[self.total:]
        for i, (p, n) in enumerate(zip(self.params, self.sizes)):
            seg = self.flat[off:off + n]
            if p.grad is None:
                seg.zero_()
                flags[i] = 0.0
            else:
                seg.copy_(p.grad.reshape(-1))
                flags[i] = 1.0
            off += n

    def unpack(self, world_size: int, average: bool = True):
        off = 0
        used = self.flat[self.total:].tolist()               # one small D2H per step
        scale = 1.0 / world_size if average else 1.0
        for p, n, u in zip(self.params, self.sizes, used):
            if u > 0:
                g = self.flat[off:off + n].view_as(p) * scale
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
            off += n

    def allreduce(self, dist=None, group=None, average: bool = True):
        """Sum (or average) the gradients of all ranks in place.  Without an initialised process group this is
        the identity (single-GPU training)."""
        if dist is None:
            import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self.pack()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.unpack(dist.get_world_size(group), average)
