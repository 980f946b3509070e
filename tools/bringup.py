"""Run every GPU parity check in its own subprocess (a trapped kernel poisons only its process).
Usage on the GPU box:  python tools/bringup.py [name ...]   -> gpurun_out/bringup.jsonl"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(name):
    import traceback
    import torch
    import gpu_checks
    t0 = time.time()
    try:
        res = gpu_checks.CHECKS[name]()
        torch.cuda.synchronize()
        print(json.dumps({"check": name, "status": "ok", "s": round(time.time() - t0, 2), **res}))
    except AssertionError as e:
        print(json.dumps({"check": name, "status": "FAIL", "s": round(time.time() - t0, 2), "detail": str(e)[:600]}))
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"check": name, "status": "ERROR", "s": round(time.time() - t0, 2),
                          "detail": (type(e).__name__ + ": " + str(e))[:600], "tb": traceback.format_exc()[-800:]}))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        run_one(sys.argv[2])
        sys.exit(0)
    import gpu_checks  # noqa: F401  (import cost only; no CUDA init)
    names = sys.argv[1:] or list(gpu_checks.CHECKS)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bringup.jsonl"), "a") as log:
        for n in names:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True,
                                   text=True, timeout=int(os.environ.get('BRINGUP_TIMEOUT', '300')), cwd=ROOT)
                lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                line = lines[-1] if lines else json.dumps({"check": n, "status": "CRASH", "rc": r.returncode,
                                                           "stderr": r.stderr[-600:]})
            except subprocess.TimeoutExpired:
                line = json.dumps({"check": n, "status": "TIMEOUT"})
            print(line, flush=True)
            log.write(line + "\n")
            log.flush()
