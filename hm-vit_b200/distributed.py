"""Data-parallel training support for the fusion module.

The reference's only parallelism is DistributedDataParallel over NCCL with find_unused_parameters=True
(opencood/tools/train_camera.py:126-131): replicas + a bucketed gradient all-reduce.  The fusion forward and
backward need no communication (scenes are independent), and this module's parameter gradients only
materialise at the end of its hand-written backward (they are pulled back from the folded-weight gradients in one
step), so there is nothing to overlap inside the module: the 2.5 M fusion parameters (10 MB fp32) travel as ONE
flat bucket in ONE all-reduce over NCCL / NVLink.

Every parameter's `.grad` IS a view into that bucket (attached at the first all-reduce), so a step costs one
`all_reduce` + one scale kernel: no per-parameter pack / unpack copies, no host synchronisation (round 1's
pack / unpack cost 6.4 ms per step at every N > 1).  Parameters that never receive a gradient (aggregate_fc, which the
reference's forward never uses, hetero_fusion.py:326) keep grad = None, as under DDP's find_unused_parameters; typed
weights of a modality absent from a batch receive exact zeros from the hand-written backward (they are part of the
folding graph) rather than None.
"""
from __future__ import annotations

from typing import List

import torch


class FlatGradAllReduce:
    """bucket = [grad_0 | grad_1 | ...] in parameter order; `p.grad` of every used parameter is a view into it."""

    def __init__(self, module: torch.nn.Module, device=None):
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.total = sum(self.sizes)
        dev = device if device is not None else (self.params[0].device if self.params else torch.device("cpu"))
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p, n in zip(self.params, self.sizes):
            self.views.append(self.flat[off:off + n].view_as(p))
            off += n
        self.attached = False

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def _attach(self):
        """Make the existing gradients views of the bucket (host-side pointer checks only: no synchronisation)."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                continue                                      # never used (aggregate_fc): stays None
            if p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        self.attached = True

    def zero_grad(self):
        """Use instead of optimizer.zero_grad(set_to_none=True): the views stay attached, one fill kernel."""
        if self.attached:
            self.flat.zero_()
        else:
            for p in self.params:
                p.grad = None

    def allreduce(self, dist=None, group=None, average: bool = True):
        """Sum (or average) the gradients of all ranks in place.  Without an initialised process group this is
        the identity (single-GPU training)."""
        if dist is None:
            import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self._attach()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.mul_(1.0 / dist.get_world_size(group))
