"""GPU: the tcgen05 (3xTF32, A in TMEM) building block of the fused decoder against an fp64 product.
fp32-class accuracy is required (the forward tolerance of the path is 1e-5 relative)."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _run(A, W, transpose):
    from miso_b200 import _lib
    lib = _lib.load()
    D = torch.empty_like(A)
    _lib.check(lib.miso_tc_selftest(A.data_ptr(), W.data_ptr(), int(transpose), D.data_ptr(), A.shape[0],
                                    _lib.stream_ptr(A.device)), "tc_selftest")
    torch.cuda.synchronize()
    return D


@pytest.mark.timeout(120)
@pytest.mark.parametrize("M", [128, 1000, 128 * 300 + 17])
@pytest.mark.parametrize("transpose", [0, 1])
def test_tc_gemm_matches_fp64(M, transpose):
    g = torch.Generator().manual_seed(M + transpose)
    A = torch.randn(M, 64, generator=g).cuda()
    W = (torch.randn(64, 64, generator=g) * 0.2).cuda()
    D = _run(A, W, transpose)
    ref = A.double() @ (W.double() if transpose else W.double().T)
    assert rel_err(D, ref) < 2e-6
    # plain fp32 matmul is the accuracy yardstick: the 3xTF32 result must be in the same class
    fp32 = A @ (W if transpose else W.T)
    assert rel_err(D, ref) < 10 * rel_err(fp32, ref) + 1e-7


@pytest.mark.timeout(120)
def test_tc_gemm_structured_values():
    """Identity / one-hot operands make layout or descriptor mistakes obvious."""
    A = torch.zeros(256, 64)
    for i in range(256):
        A[i, i % 64] = 1.0 + i
    W = torch.arange(64 * 64, dtype=torch.float32).reshape(64, 64) / 7.0
    D = _run(A.cuda(), W.cuda(), 0)
    assert rel_err(D, A.double() @ W.double().T) < 1e-6
    D = _run(A.cuda(), W.cuda(), 1)
    assert rel_err(D, A.double() @ W.double()) < 1e-6
