"""Gradients of the REFERENCE module by autograd (SURVEY 8c "grads" row) on the case of
tests/test_training_cpu.py::test_fusion_gradients_match_oracle_autograd: B=2, L=3, 256x16x24, record_len [3,2],
loss = sum(y * g_out).  Dropout is the identity (eval mode; the training configuration under test uses drop_out 0).
Stored: strided sample + norm of dL/dx, the norm and a 64-value sample of every parameter gradient, and the names of
the parameters that receive no gradient (aggregate_fc, linears of a modality absent from the batch).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_grads.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from oracle import hmvit_oracle as O  # noqa: E402


def main():
    R = ref_import.load()
    cfg = O.default_config()
    cfg["hetero_fusion_block"]["drop_out"] = 0.0
    P = O.synth_state_dict(cfg, 0)
    ref = R.HeteroFusion(cfg).eval()
    ref.load_state_dict(P, strict=True)
    x, T, m, rl, mask = O.synth_inputs(2, 3, 256, 16, 24, [3, 2], seed=11, tx=10, ty=5)
    g_out = torch.randn(2, 256, 16, 24, generator=torch.Generator().manual_seed(3))
    xg = x.clone().requires_grad_(True)
    y = ref(xg, T.clone(), m.clone(), rl.clone(), mask.clone())
    (y * g_out).sum().backward()
    out = {"dx_sample": xg.grad[:, :, ::8, ::2, ::2].contiguous().numpy(),
           "dx_norm": np.array([float(xg.grad.double().norm())]),
           "y_norm": np.array([float(y.detach().double().norm())])}
    names, norms, samples, nograd = [], [], [], []
    for name, p in ref.named_parameters():
        if p.grad is None or float(p.grad.abs().max()) == 0.0:
            nograd.append(name)
            continue
        flat = p.grad.reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        names.append(name)
        norms.append(float(flat.double().norm()))
        samples.append(flat[idx].numpy())
    out["names"] = np.array(names)
    out["norms"] = np.array(norms)
    out["samples"] = np.stack(samples)
    out["nograd"] = np.array(nograd)
    np.savez_compressed(os.path.join(HERE, "grads_c256.npz"), **out)
    print(len(names), "parameters with gradients;", len(nograd), "without:", nograd[:6], "...")


if __name__ == "__main__":
    main()
