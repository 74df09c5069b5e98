"""Host-side metrics: ``edit_distance`` with repeat merging and EOS trimming.

Mirrors utils/metrics_utils.py:8-41 (``dense_to_sparse(merge_repeated=True)`` + ``tf.edit_distance(normalize=True)``):
every row is cut at its first ``eos_id``, consecutive repeats are merged in BOTH hypothesis and truth, ``-1``
entries are dropped, and the Levenshtein distance is divided by the truth length.  Pure numpy (labels, not compute).
"""
import numpy as np


def dense_to_sequences(tensor, eos_id, merge_repeated=True):
    out = []
    for row in np.asarray(tensor):
        seq = []
        prev = None
        for t in row.tolist():
            if t == eos_id:
                break
            if t == -1:
                prev = t
                continue
            if not merge_repeated or t != prev:
                seq.append(int(t))
            prev = t
        out.append(seq)
    return out


def _levenshtein(h, t):
    d = list(range(len(t) + 1))
    for i in range(1, len(h) + 1):
        prev, d[0] = d[0], i
        for j in range(1, len(t) + 1):
            cur = min(d[j] + 1, d[j - 1] + 1, prev + (h[i - 1] != t[j - 1]))
            prev, d[j] = d[j], cur
    return d[len(t)]


def edit_distance(hypothesis, truth, eos_id, mapping=None):
    """-> float64 [B] normalised distances (tf.edit_distance: an empty truth gives inf unless the hypothesis is
    empty too, then 0)."""
    hyp, tru = np.asarray(hypothesis), np.asarray(truth)
    if mapping is not None:
        m = np.asarray(mapping)
        hyp, tru = m[hyp], m[tru]
    hs, ts = dense_to_sequences(hyp, eos_id), dense_to_sequences(tru, eos_id)
    out = np.zeros((len(hs),), np.float64)
    for i, (h, t) in enumerate(zip(hs, ts)):
        if not t:
            out[i] = 0.0 if not h else np.inf
        else:
            out[i] = _levenshtein(h, t) / len(t)
    return out


def ctc_greedy_decode(best_path, lengths, blank):
    """tf.nn.ctc_greedy_decoder(merge_repeated=True) + tf.sparse.to_dense (model_helper.py:351-353): ``best_path`` [B, T] = the
    per-frame argmax class; per utterance the first ``lengths[b]`` frames are collapsed (repeats merged, then ``blank`` removed;
    the decoder's blank is the LAST class, num_classes - 1) and the rows are padded with 0 to the longest result."""
    seqs = []
    for row, n in zip(np.asarray(best_path), np.asarray(lengths)):
        out, prev = [], None
        for c in row[:int(n)].tolist():
            if c != prev and c != blank:
                out.append(int(c))
            prev = c
        seqs.append(out)
    width = max((len(q) for q in seqs), default=0)
    dense = np.zeros((len(seqs), width), np.int32)
    for i, q in enumerate(seqs):
        dense[i, :len(q)] = q
    return dense
