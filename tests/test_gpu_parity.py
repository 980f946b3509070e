"""GPU parity suite (run on the B200 box: python -m pytest tests -m gpu).  Every test calls the CUDA
path through the C-ABI and compares with the CPU oracle / the reference's golden outputs."""
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in gpu_checks.CHECKS if n != "probe"])
def test_gpu_check(name):
    res = gpu_checks.CHECKS[name]()
    print(name, res)


def _run_with_env(name, **env_extra):
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "bringup.py"), "--one", name], capture_output=True,
                       text=True, env=env, timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert line, r.stdout[-500:] + r.stderr[-500:]
    res = json.loads(line[-1])
    assert res["status"] == "ok", res


SINGLE_CHECKS = ["attention_golden", "fusion_golden", "ragged_batch", "fusion_properties"]


@pytest.mark.parametrize("name", SINGLE_CHECKS)
def test_gpu_check_single_kernel_attention(name):
    """The default attention is the split form (csrc/attn_split.cuh); HMVIT_ATTN_SPLIT=0 selects the single fused
    warp + mask + attention kernel (csrc/attn.cuh), which stays parity-tested.  The switch is read once per process."""
    _run_with_env(name, HMVIT_ATTN_SPLIT="0")


DENSE_TC_CHECKS = ["attention_golden", "fusion_golden", "fusion_config2_scene", "ragged_batch", "fusion_properties"]


@pytest.mark.parametrize("name", DENSE_TC_CHECKS)
def test_gpu_check_tcgen05_dense_attention(name):
    """Split attention with its second launch on tcgen05 / TMEM (csrc/attn_dense_tc.cuh, HMVIT_DENSE_IMPL=tc) instead of
    the default mma.sync dense kernel.  The switch is read once per process, hence the subprocess."""
    _run_with_env(name, HMVIT_DENSE_IMPL="tc")


TC_CHECKS = ["attention_golden", "fusion_golden", "fusion_config2_scene", "ragged_batch", "fusion_properties"]


@pytest.mark.parametrize("name", TC_CHECKS)
def test_gpu_check_tcgen05_attention(name):
    """The same parity checks with the warp-specialised tcgen05 / TMEM attention kernel (attn_tc.cuh) selected
    instead of the default mma.sync one.  The switch is read once per process, hence the subprocess."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HMVIT_ATTN_IMPL="tc")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "bringup.py"), "--one", name], capture_output=True,
                       text=True, env=env, timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert line, r.stdout[-500:] + r.stderr[-500:]
    res = json.loads(line[-1])
    assert res["status"] == "ok", res
