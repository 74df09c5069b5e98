"""3xTF32 tensor-core GEMM (csrc/gemm_tf32.cu) through the C-ABI vs float64: fp32-level accuracy on the three contraction forms
of the LSTMCell matmul (las/ops.py:11-12): x W (+ b), dZ W^T and X^T dZ, with ragged M / N / K and accumulation into C."""
import ctypes as C

import numpy as np
import pytest

from tests.util import gpu


def _run(form, M, N, K, bias, accumulate, seed):
    import torch
    from phones_las_b200 import _lib
    from phones_las_b200.train import gemm_tc, split3
    rng = np.random.default_rng(seed)
    g = lambda *s: rng.standard_normal(s).astype(np.float32) * rng.uniform(0.5, 2.0)
    bia = g(N) if bias else None
    c0 = g(M, N)
    if form == "nn":      # C = A[M][K] @ W[K][N]
        a, b = g(M, K), g(K, N)
        ref = a.astype(np.float64) @ b.astype(np.float64)
    elif form == "nt":    # C = A[M][K] @ W[N][K]^T
        a, b = g(M, K), g(N, K)
        ref = a.astype(np.float64) @ b.astype(np.float64).T
    else:                 # C = A[K][M]^T @ B[K][N]
        a, b = g(K, M), g(K, N)
        ref = a.astype(np.float64).T @ b.astype(np.float64)
    if bias:
        ref = ref + bia.astype(np.float64)
    if accumulate:
        ref = ref + c0.astype(np.float64)
    ad, bd, cd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c0).cuda()
    bd_bias = torch.from_numpy(bia).cuda() if bias else None
    dev = ad.device
    if form == "nn":
        a3, seg = split3(ad.data_ptr(), M, K, K, 0, False, dev)
        b3, _ = split3(bd.data_ptr(), K, N, N, 1, True, dev)
    elif form == "nt":
        a3, seg = split3(ad.data_ptr(), M, K, K, 0, False, dev)
        b3, _ = split3(bd.data_ptr(), N, K, K, 1, False, dev)
    else:
        a3, seg = split3(ad.data_ptr(), K, M, M, 0, True, dev)
        b3, _ = split3(bd.data_ptr(), K, N, N, 1, True, dev)
    gemm_tc(a3, b3, M, N, seg, cd.data_ptr(), N, bias=bd_bias.data_ptr() if bias else None, accumulate=accumulate)
    torch.cuda.synchronize()
    got = cd.cpu().numpy().astype(np.float64)
    scale = np.abs(a).max() * np.abs(b).max() * np.sqrt(K)
    return float(np.abs(got - ref).max() / scale), a3, seg


@gpu
@pytest.mark.parametrize("form,M,N,K,bias,acc", [("nn", 300, 256, 96, True, False), ("nn", 1000, 1024, 39, True, False),
                                                  ("nt", 515, 512, 1024, False, True), ("nt", 129, 40, 260, False, False),
                                                  ("tn", 39, 1024, 2000, False, False), ("tn", 520, 384, 4500, False, True),
                                                  ("nn", 9504, 2048, 512, True, False)])
def test_tf32x3_gemm_is_fp32_accurate(form, M, N, K, bias, acc):
    err, a3, seg = _run(form, M, N, K, bias, acc, seed=M + N + K)
    # an exact-fp32 GEMM sits at ~1e-7 of max|a| max|b| sqrt(K) here, a single-TF32 product at ~3e-4; the split's own floor is
    # 3 x 2^-22 per product (rounding of the two lo parts + the dropped lo.lo term), measured 1 - 2.5e-6 on the worst element
    assert err < 5e-6, f"{form} {M}x{N}x{K}: scaled max error {err:.3e}"
    assert seg % 32 == 0 and a3.shape[1] == 3 * seg


@gpu
def test_split3_layout_and_exactness():
    """hi + lo reproduces the fp32 value to 2^-22 relative, hi and lo are TF32-representable (low 13 mantissa bits clear), the
    three segments follow the (hi, lo, hi) / (hi, hi, lo) patterns and the padding is zero."""
    import torch
    from phones_las_b200.train import split3
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((37, 50)) * 10.0 ** rng.uniform(-3, 3, (37, 50))).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    for pattern in (0, 1):
        for transpose in (False, True):
            out, seg = split3(xd.data_ptr(), 37, 50, 50, pattern, transpose, xd.device)
            o = out.cpu().numpy()
            src = x.T if transpose else x
            inner = src.shape[1]
            assert seg == 64 and o.shape == (src.shape[0], 192)
            hi, mid, last = o[:, :inner], o[:, seg:seg + inner], o[:, 2 * seg:2 * seg + inner]
            lo = mid if pattern == 0 else last
            assert np.array_equal(hi, last if pattern == 0 else mid)
            assert (o[:, inner:seg] == 0).all() and (o[:, seg + inner:2 * seg] == 0).all() and (o[:, 2 * seg + inner:] == 0).all()
            assert ((hi.view(np.uint32) & 0x1FFF) == 0).all() and ((lo.view(np.uint32) & 0x1FFF) == 0).all()
            rel = np.abs((hi.astype(np.float64) + lo.astype(np.float64)) - src) / np.abs(src)
            assert rel.max() < 2.0 ** -21
