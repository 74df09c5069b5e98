import numpy as np

from oracle import losses as olo
from phones_las_b200 import metrics


def test_edit_distance_matches_oracle_and_quirks():
    rng = np.random.default_rng(0)
    for _ in range(50):
        B, L = 4, 12
        hyp = rng.integers(2, 7, (B, L))
        tru = rng.integers(2, 7, (B, L))
        tru[:, 0] = 5  # non-empty truth
        got = metrics.edit_distance(hyp, tru, eos_id=2)
        for b in range(B):
            ref = olo.edit_distance_merge(hyp[b].tolist(), tru[b].tolist(), eos_id=2)
            assert abs(got[b] - ref) < 1e-12
    # repeats are merged in BOTH sequences; everything after the first EOS is ignored
    assert metrics.edit_distance([[3, 3, 4, 2, 9, 9]], [[3, 4, 4, 2, 1, 1]], eos_id=2)[0] == 0.0
    assert metrics.edit_distance([[3, 5, 2]], [[3, 4, 2]], eos_id=2)[0] == 0.5
    assert metrics.edit_distance([[2, 3]], [[2, 4]], eos_id=2)[0] == 0.0
    assert np.isinf(metrics.edit_distance([[3, 2]], [[2, 4]], eos_id=2)[0])
    # optional id mapping (metrics_utils.py:33-36)
    mapping = np.array([0, 1, 2, 7, 7, 5])
    assert metrics.edit_distance([[3, 2]], [[4, 2]], eos_id=2, mapping=mapping)[0] == 0.0


def test_ctc_greedy_decode_known_answer_and_oracle():
    from oracle import losses as olo
    from phones_las_b200 import metrics
    blank = 4
    path = np.array([[1, 1, 4, 1, 2, 2, 4, 4, 3], [4, 4, 0, 0, 4, 0, 3, 3, 3]])
    got = metrics.ctc_greedy_decode(path, [9, 6], blank)
    np.testing.assert_array_equal(got, [[1, 1, 2, 3], [0, 0, 0, 0]])  # second row: frames past its length are ignored; zero padding
    rng = np.random.default_rng(0)
    logits = rng.normal(size=(5, 17, 7)).astype(np.float32)
    lens = np.array([17, 3, 9, 1, 12])
    np.testing.assert_array_equal(metrics.ctc_greedy_decode(logits.argmax(-1), lens, 6), olo.ctc_greedy_decoder(logits, lens))


def test_levenshtein_core_matches_torchaudio():
    """The Levenshtein core of metrics.edit_distance (and of the oracle's) against an independent implementation
    (torchaudio.functional.edit_distance), on random label rows cut at EOS with repeats merged by hand."""
    import pytest
    taf = pytest.importorskip("torchaudio.functional")
    rng = np.random.default_rng(7)
    for _ in range(60):
        rows = []
        for _side in range(2):
            n = int(rng.integers(1, 14))
            seq = rng.integers(3, 9, n).tolist() + [2] + rng.integers(3, 9, 3).tolist()  # labels, EOS = 2, junk after EOS
            rows.append(seq)
        width = max(len(r) for r in rows)
        hyp, tru = [r + [2] * (width - len(r)) for r in rows]
        merged = []
        for r in (hyp, tru):
            cut = r[:r.index(2)]
            merged.append([x for i, x in enumerate(cut) if i == 0 or x != cut[i - 1]])
        want = taf.edit_distance(merged[0], merged[1]) / len(merged[1])
        got = metrics.edit_distance([hyp], [tru], eos_id=2)[0]
        assert got == pytest.approx(want, abs=1e-12)
        assert olo.edit_distance_merge(hyp, tru, eos_id=2) == pytest.approx(want, abs=1e-12)
