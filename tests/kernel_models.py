"""numpy models of the CUDA kernels' index math (same pass structure, same butterflies, same
table layouts) so that the algorithms are checked on CPU before they ever reach the GPU box."""
import numpy as np


def _mi(a):  # a * (-i)
    return a.imag - 1j * a.real


def _pi(a):  # a * (+i)
    return -a.imag + 1j * a.real


def bfly(v):
    """Mirrors Bfly<R>::run in csrc/frontend.cu (v: list of R complex arrays)."""
    R = len(v)
    if R == 2:
        return [v[0] + v[1], v[0] - v[1]]
    if R == 3:
        s = 0.86602540378443864676
        t1 = v[1] + v[2]
        t2 = v[0] - 0.5 * t1
        t3 = s * (v[1] - v[2])
        return [v[0] + t1, t2 + _mi(t3), t2 + _pi(t3)]
    if R == 4:
        a, b, c, d = v[0] + v[2], v[0] - v[2], v[1] + v[3], v[1] - v[3]
        return [a + c, b + _mi(d), a - c, b + _pi(d)]
    if R == 5:
        c1, c2 = 0.30901699437494742410, -0.80901699437494742410
        s1, s2 = 0.95105651629515357212, 0.58778525229247312917
        a1, a2, b1, b2 = v[1] + v[4], v[2] + v[3], v[1] - v[4], v[2] - v[3]
        p1 = v[0] + c1 * a1 + c2 * a2
        p2 = v[0] + c2 * a1 + c1 * a2
        q1 = s1 * b1 + s2 * b2
        q2 = s2 * b1 - s1 * b2
        return [v[0] + a1 + a2, p1 + _mi(q1), p2 + _mi(q2), p2 + _pi(q2), p1 + _pi(q1)]
    if R == 8:
        h = 0.70710678118654752440
        e = bfly([v[0], v[2], v[4], v[6]])
        o = bfly([v[1], v[3], v[5], v[7]])
        o[1] = h * (o[1].real + o[1].imag) + 1j * h * (o[1].imag - o[1].real)
        o[2] = _mi(o[2])
        o[3] = h * (o[3].imag - o[3].real) - 1j * h * (o[3].real + o[3].imag)
        return [e[k] + o[k] for k in range(4)] + [e[k] - o[k] for k in range(4)]
    raise ValueError(R)


def stockham_fft(z, fac, tw):
    """fft_pass<R> chain of csrc/frontend.cu.  z [n] complex64, tw [n] complex64."""
    n = len(z)
    src = z.astype(np.complex64)
    Ns = 1
    for R in fac:
        M = n // R
        twstride = n // (Ns * R)
        dst = np.zeros(n, np.complex64)
        j = np.arange(M)
        k = j % Ns
        v = []
        for t in range(R):
            x = src[j + t * M]
            if t > 0 and Ns > 1:
                x = (x * tw[k * t * twstride]).astype(np.complex64)
            v.append(x)
        y = bfly(v)
        j0 = (j - k) * R + k
        for u in range(R):
            dst[j0 + u * Ns] = y[u].astype(np.complex64)
        src = dst
        Ns *= R
    return src


def frame_power(x, tb, librosa):
    """Unpack + power of one (already windowed-in-kernel) frame, as in fe_spectral_kernel."""
    n_fft = tb["n_fft"]
    n = n_fft // 2
    xw = (x * tb["window"]).astype(np.float32)
    z = (xw[0::2] + 1j * xw[1::2]).astype(np.complex64)
    tw = (tb["tw"][:, 0] + 1j * tb["tw"][:, 1]).astype(np.complex64)
    twu = (tb["tw_unpack"][:, 0] + 1j * tb["tw_unpack"][:, 1]).astype(np.complex64)
    Z = stockham_fft(z, tb["fac"], tw)
    # bins k and n - k from the same two points of Z (k <= n / 2): X[k] = xe + w^k xo, X[n-k] = conj(xe - w^k xo); the kernel keeps
    # 2 xe and 2 xo and folds the 1/4 into the power scale
    k = np.arange(n // 2 + 1)
    zk = Z[k]
    zm = np.conj(Z[np.where(k == 0, 0, n - k)])
    xe2 = (zk + zm).astype(np.complex64)
    xo2 = _mi(zk - zm).astype(np.complex64)
    t = (twu[k] * xo2).astype(np.complex64)
    Xa, Xb = xe2 + t, xe2 - t
    scale = np.float32(0.25 if librosa else 0.25 / n_fft)
    P = np.zeros(n + 1, np.float32)
    P[k] = (Xa.real ** 2 + Xa.imag ** 2) * scale
    P[n - k] = (Xb.real ** 2 + Xb.imag ** 2) * scale
    return P


def mel_sparse(P, tb):
    """Sparse filterbank rows as fe_spectral_kernel walks them: rows are zero-padded to multiples of four weights and may reach up
    to three bins past n_fft / 2, which the kernel keeps at zero."""
    Pp = np.concatenate([P, np.zeros(3, np.float32)])
    out = np.zeros(len(tb["fb_start"]), np.float32)
    for m, (st, ln, off) in enumerate(zip(tb["fb_start"], tb["fb_len"], tb["fb_off"])):
        out[m] = np.dot(tb["fb_w"][off:off + ln], Pp[st:st + ln])
    return out


def pack_bfrag_cols(w_kn, ntile_cols_a, ntile_cols_b):
    """mma.m16n8k16 B fragments for two n-tiles (see rec.cu / decoder.cu): returns
    [K/16][32 lanes][8] with element order (A.b0.lo, A.b0.hi, A.b1.lo, A.b1.hi, B.b0.lo, ...)."""
    K = w_kn.shape[0]
    out = np.zeros((K // 16, 32, 8), w_kn.dtype)
    for lane in range(32):
        g, q = lane // 4, lane % 4
        for ti, cols in enumerate((ntile_cols_a, ntile_cols_b)):
            col = cols[g]
            for ks in range(K // 16):
                k0 = ks * 16 + 2 * q
                out[ks, lane, 4 * ti + 0] = w_kn[k0, col]
                out[ks, lane, 4 * ti + 1] = w_kn[k0 + 1, col]
                out[ks, lane, 4 * ti + 2] = w_kn[k0 + 8, col]
                out[ks, lane, 4 * ti + 3] = w_kn[k0 + 9, col]
    return out


def monotonic_attention_backward(p, prev, da, length):
    """Mirrors the luong_monotonic / bahdanau_monotonic branch of dec_att_bwd_kernel (csrc/train_dec.cu) for one utterance:
    forward recompute of cp = cumprod_excl(1 - p) (log space, clipped) and Q = cumsum(prev / clip(cp)), then two reverse running
    sums.  p [T] saved choose probabilities, prev [T] previous alignments, da [T] gradient wrt the alignments a = p cp Q.
    Returns (dscore [T] = gradient wrt the pre-sigmoid scores, dprev [T])."""
    T = len(p)
    tiny = np.finfo(np.float32).tiny
    cp, Q = np.zeros(T), np.zeros(T)
    cs = run = 0.0
    for t in range(T):
        cp[t] = np.exp(cs)
        run += prev[t] / min(max(cp[t], 1e-10), 1.0)
        Q[t] = run
        cs += np.log(min(max(1.0 - p[t], tiny), 1.0))
    ds, dprev = np.zeros(T), np.zeros(T)
    R = E = 0.0
    for t in range(T - 1, -1, -1):
        c = min(max(cp[t], 1e-10), 1.0)
        R += da[t] * p[t] * cp[t]
        dprev[t] = R / c
        dcp = da[t] * p[t] * Q[t]
        if cp[t] >= 1e-10:
            dcp -= R * prev[t] / (c * c)
        dp = da[t] * cp[t] * Q[t]
        dlogx = E
        E += dcp * cp[t]
        om = 1.0 - p[t]
        if om >= tiny:
            dp -= dlogx / om
        ds[t] = dp * p[t] * om if t < length else 0.0
    return ds, dprev
