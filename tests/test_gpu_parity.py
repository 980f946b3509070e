"""GPU parity suite (run on the B200 box: python -m pytest tests -m gpu).  Every test calls the CUDA
path through the C-ABI and compares with the CPU oracle / the reference's golden outputs."""
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in gpu_checks.CHECKS if n != "probe"])
def test_gpu_check(name):
    res = gpu_checks.CHECKS[name]()
    print(name, res)
