"""tcgen05 / TMEM / bulk-copy plumbing (csrc/umma.cuh) against a torch fp32 matmul of the same bf16 operands."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def pack_kslab(W):
    """[N,K] -> K-slab layout [K/8][N][8] (csrc/umma.cuh)."""
    N, K = W.shape
    return W.reshape(N, K // 8, 8).permute(1, 0, 2).contiguous()


@pytest.mark.parametrize("N,K", [(256, 64), (128, 128), (256, 160), (32, 16)])
def test_umma_selftest(N, K):
    from pointnerf2studio_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(N + K)
    A = torch.randn(128, K, device="cuda", generator=g).to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    D = torch.full((128, N), float("nan"), device="cuda")
    Wp = pack_kslab(W)
    _lib.check(lib.pnerf_umma_selftest(A.data_ptr(), Wp.data_ptr(), D.data_ptr(), N, K,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "selftest")
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    torch.testing.assert_close(D, ref, rtol=1e-4, atol=1e-3)
