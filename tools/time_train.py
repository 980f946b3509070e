"""Time one training step (forward with saved activations + backward) of the fusion module at the bench shape
(config 2 / config 4: 8 scenes x 5 agents x 256x48x176) and break the step down per kernel family with CUDA
events (a timing proxy around hmvit_b200.ops).  Usage on the GPU box: python tools/time_train.py [batch]"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hmvit_loader  # noqa: E402
from oracle import hmvit_oracle as O  # noqa: E402
import bench  # noqa: E402


class TimedOps:
    """forwards every call to hmvit_b200.ops, bracketing it with CUDA events"""

    def __init__(self, ops):
        self._ops, self.events = ops, collections.defaultdict(list)

    def __getattr__(self, name):
        fn = getattr(self._ops, name)
        if not callable(fn):
            return fn

        def wrapped(*a, **k):
            tag = name
            if name == "rowgemm":
                tag = f"rowgemm[{a[0]}]"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            self.events[tag].append((e0, e1))
            return r
        return wrapped

    def summary(self):
        torch.cuda.synchronize()
        return {k: (len(v), round(sum(a.elapsed_time(b) for a, b in v), 3)) for k, v in self.events.items()}


def main():
    Bq = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0")
    pkg = hmvit_loader.load()
    cfg = O.default_config()
    cfg["hetero_fusion_block"]["drop_out"] = 0.0
    net = pkg.HeteroFusion(cfg).train()
    net.load_state_dict(O.synth_state_dict(cfg, 0))
    net = net.to(dev)
    x, T, mode, rl, mask = bench.make_inputs(1236, Bq)
    inp = [t.to(dev) for t in (T, mode, rl, mask)]
    xd = x.to(dev).requires_grad_(True)
    g = torch.randn(Bq, 256, 48, 176, device=dev)

    def step(ops):
        for p in net.parameters():
            p.grad = None
        y = pkg.training.fusion_train(ops, net.hetero_fusion_block, net, xd, *inp, num_iters=net.num_iters)
        (y * g).sum().backward()

    for _ in range(2):
        step(pkg.ops)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        step(pkg.ops)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t = TimedOps(pkg.ops)
    step(t)
    s = t.summary()
    print(json.dumps({"batch": Bq, "train_step_ms": round(ms, 3), "scenes_per_s": round(Bq / ms * 1e3, 1),
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2),
                      "ops(count, ms)": dict(sorted(s.items(), key=lambda kv: -kv[1][1]))}))


if __name__ == "__main__":
    main()
