"""tcgen05 plumbing self-test: D = A x B^T through fpc_selftest_umma (the same descriptor / TMEM
helpers the bf16 predictor uses) against a plain PyTorch fp32 reference of the same product."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(16, 16), (64, 32), (128, 64), (256, 64), (256, 192), (48, 128), (32, 416), (64, 512)])
def test_umma_matches_fp32_reference(N, K):
    import torch
    import fpc_native as NV
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + K)
    # small integers: every product and partial sum is exact in fp32 -> the result must be EXACT,
    # so any layout / descriptor mistake shows up as a hard mismatch, not as "noise"
    A = torch.randint(-3, 4, (128, K), generator=g, device="cuda").to(torch.bfloat16)
    B = torch.randint(-3, 4, (N, K), generator=g, device="cuda").to(torch.bfloat16)
    D = torch.full((128, N), float("nan"), device="cuda")
    NV.check(NV.lib().fpc_selftest_umma(A.data_ptr(), B.data_ptr(), N, K, D.data_ptr(), NV.current_stream()), "selftest")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    bad = (D != ref).nonzero()
    assert bad.numel() == 0, "first mismatch at %s: got %r want %r (%d wrong)" % (
        bad[0].tolist(), D[tuple(bad[0])].item(), ref[tuple(bad[0])].item(), bad.shape[0])
    # random bf16 values: fp32 accumulation, order-of-summation tolerance only
    A = torch.randn((128, K), generator=g, device="cuda").to(torch.bfloat16)
    B = torch.randn((N, K), generator=g, device="cuda").to(torch.bfloat16)
    NV.check(NV.lib().fpc_selftest_umma(A.data_ptr(), B.data_ptr(), N, K, D.data_ptr(), NV.current_stream()), "selftest")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    assert (D - ref).abs().max().item() <= 1e-4 * K ** 0.5 + 1e-5
