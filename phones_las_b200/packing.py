"""Re-layout of TF-checkpoint weights into the device layouts the kernels consume (numpy, host).

TF's LSTMCell stores one ``kernel [din+U, 4U]`` per cell with rows ``[W_x ; W_h]`` and gate
column blocks ``[i | j | f | o]`` (tf.nn.rnn_cell.LSTMCell, reference las/ops.py:11-12).  The
kernels want every hidden unit's four gates next to each other (so one thread can apply the gate
math) and, for the tensor-core paths, pre-packed mma.m16n8k16 B fragments.  Layout contracts are
documented in include/plas.h and DESIGN.md; tests/test_host_packing.py checks them against a
numpy model of the fragment semantics.
"""
import numpy as np

GATES = 4


def unit_major_cols(U):
    """TF column index for packed column (unit, gate): result[4*u + g] = g*U + u."""
    u = np.arange(U)
    return (np.arange(GATES)[None, :] * U + u[:, None]).reshape(-1)


def pack_inproj(kernels, biases, din, U, k_pad):
    """Input-projection operand for K2: Wt [ndir*4U, k_pad] (row = dir*4U + 4*unit + gate, K-major,
    zero padded along K) and bias [ndir*4U]."""
    cols = unit_major_cols(U)
    wt = np.zeros((len(kernels) * 4 * U, k_pad), np.float32)
    bs = np.zeros((len(kernels) * 4 * U,), np.float32)
    for d, (k, b) in enumerate(zip(kernels, biases)):
        wx = np.asarray(k[:din], np.float32)[:, cols]  # [din, 4U] packed column order
        wt[d * 4 * U:(d + 1) * 4 * U, :din] = wx.T
        bs[d * 4 * U:(d + 1) * 4 * U] = np.asarray(b, np.float32)[cols]
    return wt, bs


def pack_rec_f32(kernels, din, U, upc):
    """Recurrent weights for rec_f32_kernel: [ndir][U/upc][U(k)][4*upc], col = 4*unit_local + gate."""
    G = U // upc
    out = np.zeros((len(kernels), G, U, 4 * upc), np.float32)
    for d, k in enumerate(kernels):
        wh = np.asarray(k[din:], np.float32)  # [U, 4U]
        for ci in range(G):
            units = ci * upc + np.arange(upc)
            cols = (np.arange(GATES)[None, :] * U + units[:, None]).reshape(-1)
            out[d, ci] = wh[:, cols]
    return out


def bfrag_pack(w_kn, cols_a, cols_b):
    """mma.m16n8k16 '.col' B fragments of two n-tiles.  w_kn [K, *]; cols_a/cols_b: the 8 source
    columns of n-tile A / B.  Returns [K/16][32 lanes][8]: lane (g=lane/4, q=lane%4) holds, per tile,
    b0 = (W[k0+2q][n=g], W[k0+2q+1][g]) and b1 = (W[k0+2q+8][g], W[k0+2q+9][g])."""
    K = w_kn.shape[0]
    ks = np.arange(K // 16)[:, None, None]
    lane = np.arange(32)[None, :, None]
    g, q = lane // 4, lane % 4
    koff = np.array([0, 1, 8, 9])[None, None, :]
    krow = ks * 16 + 2 * q + koff                       # [KS,32,4]
    out = np.zeros((K // 16, 32, 8), w_kn.dtype)
    ca = np.asarray(cols_a)[g]                           # [1,32,1]
    cb = np.asarray(cols_b)[g]
    out[:, :, 0:4] = w_kn[krow, np.broadcast_to(ca, krow.shape)]
    out[:, :, 4:8] = w_kn[krow, np.broadcast_to(cb, krow.shape)]
    return out


def _tile_cols(unit0, U):
    """Source (TF) columns of the (i,j) tile and the (f,o) tile for units unit0..unit0+3."""
    units = unit0 + np.arange(8) // 2
    par = np.arange(8) % 2
    return par * U + units, (2 + par) * U + units


def pack_rec_bf16(kernels, din, U):
    """Recurrent weights for rec_bf16_kernel: [ndir][U/32][8 warps][U/16][32][8] (float32 values,
    cast to bf16 on upload).  Warp w of slice ci owns units ci*32 + 4w .. +3."""
    G = U // 32
    out = np.zeros((len(kernels), G, 8, U // 16, 32, 8), np.float32)
    for d, k in enumerate(kernels):
        wh = np.asarray(k[din:], np.float32)
        for ci in range(G):
            for w in range(8):
                ca, cb = _tile_cols(ci * 32 + 4 * w, U)
                out[d, ci, w] = bfrag_pack(wh, ca, cb)
    return out


def pack_rec_tc(kernels, din, U):
    """Recurrent weights for rec_tc_kernel (TMEM-resident A operand): [ndir][U/32][128][U] (K-major).  Row (TMEM
    lane) m = 32*q + 8*gate + u8 of CTA ci holds W_h[:, gate*U + ci*32 + 8*q + u8]: inside every 32-lane quadrant
    the four gates of a unit sit 8 lanes apart, which is what the 16x256b TMEM load hands to one thread."""
    out = np.zeros((len(kernels), U // 32, 128, U), np.float32)
    q, gate, u8 = np.meshgrid(np.arange(4), np.arange(4), np.arange(8), indexing="ij")
    m = (32 * q + 8 * gate + u8).reshape(-1)
    ul = (8 * q + u8).reshape(-1)
    g = gate.reshape(-1)
    for d, k in enumerate(kernels):
        wh = np.asarray(k[din:], np.float32)  # [U, 4U]
        for ci in range(U // 32):
            out[d, ci, m] = wh[:, g * U + ci * 32 + ul].T
    return out


def pack_cell_f32(w_rows, Ud):
    """Decoder LSTM rows (non-embedding part) for decoder_kernel<float>: [Ud/4][K][16], col = 4*ul+gate."""
    K = w_rows.shape[0]
    nsl = Ud // 4
    out = np.zeros((nsl, K, 16), np.float32)
    for s in range(nsl):
        units = 4 * s + np.arange(4)
        cols = (np.arange(GATES)[None, :] * Ud + units[:, None]).reshape(-1)
        out[s] = w_rows[:, cols]
    return out


def pack_cell_bf16(w_rows, Ud):
    """Decoder LSTM rows for decoder_kernel<bf16>: [Ud/4][K/16][32][8] B fragments."""
    K = w_rows.shape[0]
    nsl = Ud // 4
    out = np.zeros((nsl, K // 16, 32, 8), np.float32)
    for s in range(nsl):
        ca, cb = _tile_cols(4 * s, Ud)
        out[s] = bfrag_pack(np.asarray(w_rows, np.float32), ca, cb)
    return out


def swizzle128_tiles(w_nk):
    """UMMA/TMA K-major SWIZZLE_128B image of a [16, K] operand: per 64-wide k block a 16 x 128-byte
    tile whose 16-byte chunk c of row r is stored at chunk position c ^ (r % 8).  Returns [K/64][1024]."""
    n, K = w_nk.shape
    assert n == 16 and K % 64 == 0
    t = np.asarray(w_nk, np.float32).reshape(16, K // 64, 8, 8)  # [row][k block][chunk][element]
    out = np.empty((K // 64, 16, 8, 8), np.float32)
    for r in range(16):
        for c in range(8):
            out[:, r, c ^ (r % 8), :] = t[r, :, c, :]
    return out.reshape(K // 64, 1024)


def pack_cell_tc(w_rows, Ud):
    """Decoder LSTM rows for decoder_tc_kernel: [Ud/4][K/64][1024]; tile row = 4*unit_local + gate."""
    K = w_rows.shape[0]
    nsl = Ud // 4
    out = np.zeros((nsl, K // 64, 1024), np.float32)
    for s in range(nsl):
        units = 4 * s + np.arange(4)
        cols = (np.arange(GATES)[None, :] * Ud + units[:, None]).reshape(-1)
        out[s] = swizzle128_tiles(np.asarray(w_rows, np.float32)[:, cols].T)
    return out


def pack_query_tc(wq, Ud):
    """bahdanau query_layer kernel [Ud(k), Ud(out)] for decoder_tc_kernel: [Ud/16][Ud/64][1024]."""
    out = np.zeros((Ud // 16, Ud // 64, 1024), np.float32)
    for s in range(Ud // 16):
        out[s] = swizzle128_tiles(np.asarray(wq, np.float32)[:, 16 * s:16 * s + 16].T)
    return out


def pack_unit_major(mat_or_vec, U):
    """Permute the last axis from TF gate-block order to (unit, gate) order."""
    return np.asarray(mat_or_vec, np.float32)[..., unit_major_cols(U)]
