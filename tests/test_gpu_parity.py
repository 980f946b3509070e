"""GPU parity suite (run on the B200 box: python -m pytest tests -m gpu).  Every test calls the CUDA
path through the C-ABI and compares with the CPU oracle / the reference's golden outputs."""
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(gpu_checks.CHECKS))
def test_gpu_check(name):
    res = gpu_checks.CHECKS[name]()
    print(name, res)


CROSS_CHECKS = ["attention_golden", "fusion_golden", "fusion_config2_scene", "ragged_batch", "fusion_properties"]


@pytest.mark.parametrize("impl", ["split", "single"])
@pytest.mark.parametrize("name", CROSS_CHECKS)
def test_gpu_check_other_attention_impl(name, impl):
    """The default attention is the fused persistent tcgen05 kernel (csrc/attn_fused.cuh).  The split form
    (csrc/attn_split.cuh) and the single mma.sync kernel (csrc/attn.cuh) are independently written implementations of
    the same contract, selected explicitly through HmvitAttnArgs.impl / HmvitFusionArgs.attn_impl: the same parity
    checks must hold for them."""
    p = gpu_checks.pkg()
    p.HeteroFusionBlock.attn_impl = p.HeteroAttention.attn_impl = impl      # class-level default read by every instance
    try:
        res = gpu_checks.CHECKS[name]()
    finally:
        p.HeteroFusionBlock.attn_impl = p.HeteroAttention.attn_impl = None
    print(name, impl, res)
