"""Decoder (HeteroDecoder on the ego's fused feature) at the bench shape: parity numbers of the GPU checks and the time of
hmvit_decoder_forward for 8 scenes of 256x48x176 (4 x conv3x3 256->256 + heads = 39.9 GF per scene)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import hmvit_loader
from oracle import hmvit_oracle as O
import gpu_checks

dev = torch.device("cuda:0")
pkg = hmvit_loader.load()
res = {"decoder_vs_oracle": gpu_checks.check_decoder_vs_oracle(), "decoder_logits_golden": gpu_checks.check_decoder_logits_golden()}
B, H, W = 8, 48, 176
dec = pkg.HeteroDecoder({"input_dim": 256, "num_layer": 2, "num_ch_dec": [256, 256], "anchor_number": 2}).eval()
dec.load_state_dict(O.synth_decoder_state_dict(1), strict=True)
dec = dec.to(dev)
x = torch.randn(B, 256, H, W, device=dev)
mode = torch.tensor([[b & 1] for b in range(B)], dtype=torch.int32, device=dev)
with torch.no_grad():
    for _ in range(5):
        dec(x, mode, use_upsample=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dec(x, mode, use_upsample=False)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
gf = 4 * 2 * 9 * 256 * 256 * H * W * B / 1e9 + 2 * 16 * 256 * H * W * B / 1e9
res["decoder_ms_per_8_scenes"] = round(ms, 4)
res["decoder_tflops"] = round(gf / ms, 1)
print(json.dumps(res))
