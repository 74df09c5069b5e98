"""K2 parity: tcgen05 bf16 GEMM and exact-fp32 GEMM (through the C-ABI) vs numpy/torch matmul."""
import numpy as np
import pytest

from tests.util import gpu, to_np, scaled_err


def _run(precision, M, K, N, lda=None, with_bias=True, seed=0):
    import torch
    from phones_las_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(seed)
    lda = lda or K
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    a = (torch.randn((M, lda), generator=g) * 0.5).to(dt).cuda()
    w = (torch.randn((N, K), generator=g) * 0.1).to(dt).cuda()
    bias = torch.randn((N,), generator=g).cuda() if with_bias else None
    c = torch.full((M, N), float("nan"), dtype=dt).cuda()
    fn = L.plas_gemm_bf16 if precision == "bf16" else L.plas_gemm_f32
    _lib.check(fn(_lib.ptr(a), M, K, lda, _lib.ptr(w), N, K, _lib.ptr(bias) if with_bias else None,
                  _lib.ptr(c), N, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = a[:, :K].double().cpu() @ w.double().cpu().T
    if with_bias:
        ref = ref + bias.double().cpu()
    return to_np(c), ref.numpy()


@gpu
@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (128, 128, 256), (300, 192, 256), (1000, 1024, 512),
                                   (77, 128, 128), (4096, 2048, 4096), (513, 80, 384)])
def test_gemm_bf16_tcgen05(M, K, N):
    lda = K if K % 8 == 0 else (K + 7) // 8 * 8
    c, ref = _run("bf16", M, K, N, lda=max(lda, K))
    assert np.isfinite(c).all()
    # output is bf16: one rounding of an fp32-accumulated value
    err = np.abs(c - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() <= 2.0 ** -8 + 1e-3, f"max rel err {err.max():.3e}"
    assert np.linalg.norm(c - ref) / np.linalg.norm(ref) <= 3e-3


@gpu
def test_gemm_bf16_k_tail_zero_fill():
    # K = 80 (not a multiple of the 64-wide TMA box): out-of-bounds columns must read as zero
    c, ref = _run("bf16", 256, 80, 128, lda=128)
    assert np.linalg.norm(c - ref) / np.linalg.norm(ref) <= 3e-3


@gpu
@pytest.mark.parametrize("M,K,N", [(64, 16, 64), (130, 39, 1024), (257, 295, 96), (1000, 768, 2048)])
def test_gemm_f32_exact(M, K, N):
    c, ref = _run("fp32", M, K, N)
    assert scaled_err(c, ref) <= 2e-6
