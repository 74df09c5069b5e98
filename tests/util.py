"""Shared helpers for parity tests: tolerance rules (DESIGN.md "parity bar") and small builders."""
import numpy as np
import pytest

try:
    import torch
    HAS_CUDA = torch.cuda.is_available()
except Exception:  # pragma: no cover
    HAS_CUDA = False

gpu = pytest.mark.gpu


def to_np(t):
    import torch
    return t.detach().to(torch.float32).cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def scaled_err(a, b):
    """max |a-b| relative to the reference tensor's scale (max |b|)."""
    a, b = to_np(a).astype(np.float64), to_np(b).astype(np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) if b.size else 0.0


def fro_err(a, b):
    a, b = to_np(a).astype(np.float64), to_np(b).astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)) if b.size else 0.0


def assert_parity(a, b, precision, what="", bf16_fro=1e-3):
    """fp32: max error <= 1e-5 of the tensor scale.  bf16 (oracle emulates the same rounding
    points): relative Frobenius error <= ``bf16_fro`` (default 1e-3, north_star's figure), and no
    element off by more than one bf16 ulp of the tensor scale (2^-7) -- a single rounding flip of a
    value in [0.5,1) is 2^-8.  Deep recurrent stacks pass bf16_fro=2e-3 (half a bf16 epsilon): a
    1-ulp flip of one h feeds back through W_hh and flips a few percent of later roundings, so two
    correct bf16 implementations with different fp32 accumulation order drift apart by ~1.2e-3 after
    four layers (measured; the bf16-vs-fp32 oracle gap itself is 3e-3)."""
    a, b = to_np(a), to_np(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if precision == "fp32":
        e = scaled_err(a, b)
        assert e <= 1e-5, f"{what}: fp32 scaled max error {e:.3e} > 1e-5"
    else:
        f, e = fro_err(a, b), scaled_err(a, b)
        assert f <= bf16_fro, f"{what}: bf16 relative Frobenius error {f:.3e} > {bf16_fro:g}"
        assert e <= 2.0 ** -7, f"{what}: bf16 scaled max error {e:.3e} > 2^-7"


def top2_margin(logits):
    s = np.sort(to_np(logits), axis=-1)
    return float((s[..., -1] - s[..., -2]).min()) if s.size else np.inf
