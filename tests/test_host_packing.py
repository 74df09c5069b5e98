"""Weight re-layout (TF checkpoint -> kernel layouts) checked against a numpy model of the
mma.m16n8k16 fragment semantics and of the kernels' column conventions."""
import numpy as np

from phones_las_b200 import packing


def _b_from_frags(frag_tile):
    """Rebuild B[16][8] of one k-step from per-lane (b0.lo,b0.hi,b1.lo,b1.hi) (PTX m16n8k16 .col B)."""
    Bm = np.zeros((16, 8), frag_tile.dtype)
    for lane in range(32):
        g, q = lane // 4, lane % 4
        Bm[2 * q, g], Bm[2 * q + 1, g], Bm[2 * q + 8, g], Bm[2 * q + 9, g] = frag_tile[lane]
    return Bm


def _tf_lstm_z(x_h, kernel, bias):
    return x_h @ kernel + bias


def test_unit_major_cols():
    U = 6
    cols = packing.unit_major_cols(U)
    assert cols[4 * 3 + 2] == 2 * U + 3  # unit 3, gate f


def test_pack_inproj_matches_tf_projection():
    rng = np.random.default_rng(0)
    din, U = 5, 8
    ks = [rng.standard_normal((din + U, 4 * U)).astype(np.float32) for _ in range(2)]
    bs = [rng.standard_normal(4 * U).astype(np.float32) for _ in range(2)]
    wt, b = packing.pack_inproj(ks, bs, din, U, k_pad=64)
    x = rng.standard_normal((3, din)).astype(np.float32)
    xp = np.pad(x, ((0, 0), (0, 64 - din)))
    out = xp @ wt.T + b
    for d in range(2):
        ref = x @ ks[d][:din] + bs[d]  # TF order [i|j|f|o] blocks
        for u in range(U):
            for g in range(4):
                np.testing.assert_allclose(out[:, d * 4 * U + 4 * u + g], ref[:, g * U + u], rtol=1e-5)


def test_pack_rec_f32_layout():
    rng = np.random.default_rng(1)
    din, U, upc = 3, 16, 8
    k = rng.standard_normal((din + U, 4 * U)).astype(np.float32)
    p = packing.pack_rec_f32([k], din, U, upc)
    assert p.shape == (1, 2, U, 4 * upc)
    h = rng.standard_normal((2, U)).astype(np.float32)
    ref = h @ k[din:]
    for ci in range(2):
        z = h @ p[0, ci]
        for ul in range(upc):
            for g in range(4):
                np.testing.assert_allclose(z[:, 4 * ul + g], ref[:, g * U + ci * upc + ul], rtol=1e-5)


def test_pack_rec_bf16_fragments():
    rng = np.random.default_rng(2)
    din, U = 4, 64
    k = rng.standard_normal((din + U, 4 * U)).astype(np.float32)
    p = packing.pack_rec_bf16([k], din, U)
    assert p.shape == (1, 2, 8, U // 16, 32, 8)
    h = rng.standard_normal((16, U)).astype(np.float32)
    ref = h @ k[din:]
    for ci in (0, 1):
        for w in (0, 5):
            accA = np.zeros((16, 8))
            accB = np.zeros((16, 8))
            for ks in range(U // 16):
                A = h[:, ks * 16:(ks + 1) * 16]
                accA += A @ _b_from_frags(p[0, ci, w, ks, :, 0:4])
                accB += A @ _b_from_frags(p[0, ci, w, ks, :, 4:8])
            for q in range(4):  # kernel: thread (g,q) reads acc[2q], acc[2q+1] as (i,j) / (f,o)
                unit = ci * 32 + 4 * w + q
                np.testing.assert_allclose(accA[:, 2 * q], ref[:, 0 * U + unit], rtol=1e-4, atol=1e-5)
                np.testing.assert_allclose(accA[:, 2 * q + 1], ref[:, 1 * U + unit], rtol=1e-4, atol=1e-5)
                np.testing.assert_allclose(accB[:, 2 * q], ref[:, 2 * U + unit], rtol=1e-4, atol=1e-5)
                np.testing.assert_allclose(accB[:, 2 * q + 1], ref[:, 3 * U + unit], rtol=1e-4, atol=1e-5)


def test_pack_cell_layouts():
    rng = np.random.default_rng(3)
    K, Ud = 32, 16
    rows = rng.standard_normal((K, 4 * Ud)).astype(np.float32)
    x = rng.standard_normal((5, K)).astype(np.float32)
    ref = x @ rows
    pf = packing.pack_cell_f32(rows, Ud)
    pb = packing.pack_cell_bf16(rows, Ud)
    assert pf.shape == (Ud // 4, K, 16) and pb.shape == (Ud // 4, K // 16, 32, 8)
    for s in range(Ud // 4):
        z = x @ pf[s]
        part = np.zeros((5, 16))
        for ks in range(K // 16):
            A = x[:, ks * 16:(ks + 1) * 16]
            part[:, 0:8] += A @ _b_from_frags(pb[s, ks, :, 0:4])
            part[:, 8:16] += A @ _b_from_frags(pb[s, ks, :, 4:8])
        for ul in range(4):
            for g in range(4):
                u = 4 * s + ul
                np.testing.assert_allclose(z[:, 4 * ul + g], ref[:, g * Ud + u], rtol=1e-4, atol=1e-5)
                col = (g >> 1) * 8 + 2 * ul + (g & 1)  # gate_col<bf16> in decoder.cu
                np.testing.assert_allclose(part[:, col], ref[:, g * Ud + u], rtol=1e-4, atol=1e-5)
    pu = packing.pack_unit_major(rows, Ud)
    np.testing.assert_array_equal(pu[:, 4 * 5 + 3], rows[:, 3 * Ud + 5])


def test_swizzle128_tiles_roundtrip():
    """Element (row r, k) of a [16, K] operand must sit at byte kb*2048 + r*128 + ((c ^ (r & 7)) << 4) + (k & 7)*2."""
    rng = np.random.default_rng(0)
    w = rng.standard_normal((16, 192)).astype(np.float32)
    t = packing.swizzle128_tiles(w)  # [3][1024]
    for r in range(16):
        for k in range(192):
            kb, c, e = k // 64, (k % 64) // 8, k % 8
            pos = r * 64 + ((c ^ (r & 7)) * 8) + e
            assert t[kb, pos] == w[r, k]


def test_pack_cell_tc_and_query_tc():
    rng = np.random.default_rng(1)
    Ud, K = 64, 128
    rows = rng.standard_normal((K, 4 * Ud)).astype(np.float32)
    t = packing.pack_cell_tc(rows, Ud)
    assert t.shape == (Ud // 4, K // 64, 1024)
    s, ul, gate, k = 5, 2, 3, 77   # tile row 4*ul+gate of slice s holds TF column gate*Ud + 4*s+ul
    r = 4 * ul + gate
    kb, c, e = k // 64, (k % 64) // 8, k % 8
    assert t[s, kb, r * 64 + ((c ^ (r & 7)) * 8) + e] == rows[k, gate * Ud + 4 * s + ul]
    wq = rng.standard_normal((Ud, Ud)).astype(np.float32)
    tq = packing.pack_query_tc(wq, Ud)
    s, r, k = 2, 9, 40
    kb, c, e = k // 64, (k % 64) // 8, k % 8
    assert tq[s, kb, r * 64 + ((c ^ (r & 7)) * 8) + e] == wq[k, 16 * s + r]


def test_pack_rec_tc_rows():
    rng = np.random.default_rng(2)
    U, din = 64, 5
    kernels = [rng.standard_normal((din + U, 4 * U)).astype(np.float32) for _ in range(2)]
    t = packing.pack_rec_tc(kernels, din, U)
    assert t.shape == (2, U // 32, 128, U)
    d, ci, ul, gate, k = 1, 1, 13, 2, 33  # unit 13 of the CTA: quadrant 1, u8 = 5
    assert t[d, ci, 32 * (ul // 8) + 8 * gate + ul % 8, k] == kernels[d][din + k, gate * U + ci * 32 + ul]


def test_recurrence_cluster_budget_policy():
    """listener.rec_cluster_budget: the serving loops' 64-SM budget becomes 4 clusters of 16 CTAs at U = 512 (8 of 8 at U = 256) and
    is dropped when the batch would then need more than two 16-row groups per cluster."""
    from phones_las_b200.listener import rec_cluster_budget
    assert rec_cluster_budget(64, 512, 2, 0) == 0          # no budget: every placeable cluster
    assert rec_cluster_budget(64, 512, 2, 64) == 4         # c2: 2 clusters per direction x 2 groups x 16 rows
    assert rec_cluster_budget(16, 512, 2, 64) == 4
    assert rec_cluster_budget(65, 512, 2, 64) == 0         # 17 rows per group: budget dropped
    assert rec_cluster_budget(128, 512, 2, 64) == 0        # c4 on one GPU: 3 x 15 rows on 6 clusters instead
    assert rec_cluster_budget(64, 256, 2, 64) == 8
    assert rec_cluster_budget(64, 512, 1, 64) == 4         # unidirectional: 4 clusters, 1 group of 16 each
    assert rec_cluster_budget(200, 512, 1, 64) == 0
    assert rec_cluster_budget(8, 512, 2, 16) == 2          # never fewer clusters than directions
