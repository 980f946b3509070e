"""BASELINE config 3 side figure alone (bench.py::config3_variant) + the two f-3 GPU checks."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import hmvit_loader
import bench
import gpu_checks

pkg = hmvit_loader.load()
res = {"encoders_golden": gpu_checks.check_encoders_golden(), "config3_golden": gpu_checks.check_config3_golden()}
res["config3_variant"] = bench.config3_variant(pkg, torch.device("cuda:0"), scenes=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
print(json.dumps(res))
