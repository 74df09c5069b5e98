"""Oracle: losses, prediction heads and the optimiser step (numpy; test infrastructure only).

Restates model_helper.py:20-146 (compute_loss, sequence_loss_sigmoid, compute_loss_sigmoid,
compute_log_probs_loss), :347-358 (CTC head), :403-417 (L2 + per-tensor clip + Adam) and
utils/training_helper.py:17-27 (transform_binf_to_phones), plus the TF 1.15.2 ops they call
(tf.contrib.seq2seq.sequence_loss, tf.nn.ctc_loss_v2, tf.nn.ctc_greedy_decoder,
tf.clip_by_norm, tf.train.AdamOptimizer).  Parity against TF itself is UNPINNED
(oracle/__init__.py); CTC is cross-checked against torch.nn.functional.ctc_loss.
"""
import numpy as np

F32 = np.float32


def log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    y = x - m
    return y - np.log(np.exp(y).sum(axis=-1, keepdims=True))


def sequence_loss(logits, targets, weights):
    """tf.contrib.seq2seq.sequence_loss defaults: sum(CE*w) / (sum(w) + 1e-12)."""
    lp = log_softmax(logits.astype(np.float64))
    B, T, V = logits.shape
    ce = -np.take_along_axis(lp, targets[..., None].astype(np.int64), axis=-1)[..., 0]
    return float((ce * weights).sum() / (weights.sum() + 1e-12))


def sequence_mask(lengths, maxlen):
    return (np.arange(maxlen)[None, :] < np.asarray(lengths)[:, None]).astype(np.float64)


def compute_loss(logits, targets, final_sequence_length, target_sequence_length, mode, eos_id):
    """model_helper.py:20-78."""
    if mode == "train":
        T = targets.shape[1]
        if logits.shape[1] < T:
            logits = np.pad(logits, [(0, 0), (0, T - logits.shape[1]), (0, 0)])
        w = sequence_mask(target_sequence_length, T)
        return sequence_loss(logits, targets, w)
    max_ts = int(np.max(target_sequence_length))
    max_fs = int(np.max(final_sequence_length))
    L = max(max_ts, max_fs)
    logits = logits[:, :max_fs]
    if targets.shape[1] < L:
        targets = np.pad(targets, [(0, 0), (0, L - targets.shape[1])], constant_values=eos_id)
    if logits.shape[1] < L:
        logits = np.pad(logits, [(0, 0), (0, L - logits.shape[1]), (0, 0)])
    seq_len = np.maximum(np.asarray(target_sequence_length), np.asarray(final_sequence_length))
    w = sequence_mask(seq_len, L)
    # tf.sequence_mask(maxlen=L) vs targets possibly longer than L: the reference pads only
    return sequence_loss(logits[:, :L], targets[:, :L], w)


def sigmoid_ce(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|))."""
    x = logits.astype(np.float64)
    z = labels.astype(np.float64)
    return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))


def sequence_loss_sigmoid(logits, targets, weights):
    """model_helper.py:81-95."""
    n = logits.shape[2]
    ce = sigmoid_ce(logits.reshape(-1, n), targets.reshape(-1, n)).mean(axis=1)
    w = weights.reshape(-1)
    return float((ce * w).sum() / (w.sum() + 1e-12))


def compute_loss_sigmoid_train(logits, targets_binf, target_sequence_length):
    """model_helper.py:98-105 (TRAIN branch)."""
    w = sequence_mask(target_sequence_length, logits.shape[1])
    return sequence_loss_sigmoid(logits, targets_binf, w)


def compute_log_probs_loss(outputs):
    """model_helper.py:132-146."""
    n = outputs.shape[-1] // 2
    o = outputs.astype(np.float64)
    l1, l0 = o[..., :n], o[..., n:2 * n]
    c = -(l1 + l0) / 2
    loss = np.abs((np.exp(l1 + c) + np.exp(l0 + c)) / np.exp(c) - 1)
    loss = loss + np.maximum(l1, 0) + np.maximum(l0, 0)
    return float(loss.mean())


def transform_binf_to_phones(outputs, binf_to_ipa):
    """utils/training_helper.py:17-27."""
    n = binf_to_ipa.shape[0]
    return outputs[..., :n] @ binf_to_ipa + outputs[..., n:2 * n] @ (1 - binf_to_ipa)


def ctc_loss(logits, labels, label_length, logit_length, blank=0):
    """tf.nn.ctc_loss_v2 dense-label path (model_helper.py:355-356): blank index 0,
    loss[b] = -log p(labels[b,:label_length[b]] | logits[b,:logit_length[b]]).  Returns [B]."""
    B = logits.shape[0]
    out = np.zeros((B,), np.float64)
    for b in range(B):
        T = int(logit_length[b])
        L = int(label_length[b])
        lp = log_softmax(logits[b, :T].astype(np.float64))
        ext = np.full((2 * L + 1,), blank, np.int64)
        ext[1::2] = labels[b, :L]
        S = 2 * L + 1
        alpha = np.full((S,), -np.inf)
        if T == 0:
            out[b] = 0.0 if L == 0 else np.inf
            continue
        alpha[0] = lp[0, ext[0]]
        if S > 1:
            alpha[1] = lp[0, ext[1]]
        for t in range(1, T):
            prev = alpha
            a1 = np.concatenate([[-np.inf], prev[:-1]])
            a2 = np.concatenate([[-np.inf, -np.inf], prev[:-2]])
            can_skip = np.zeros((S,), bool)
            can_skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
            a2 = np.where(can_skip, a2, -np.inf)
            stacked = np.stack([prev, a1, a2])
            m = stacked.max(axis=0)
            safe_m = np.where(np.isfinite(m), m, 0.0)
            with np.errstate(divide="ignore"):
                alpha = safe_m + np.log(np.exp(stacked - safe_m).sum(axis=0))
            alpha = np.where(np.isfinite(m), alpha, -np.inf) + lp[t, ext]
        tail = alpha[-1] if S == 1 else np.logaddexp(alpha[-1], alpha[-2])
        out[b] = -tail
    return out


def ctc_greedy_decode(logits, seq_len):
    """tf.nn.ctc_greedy_decoder: argmax per frame, merge repeats, drop blank = LAST class."""
    res = []
    blank = logits.shape[-1] - 1
    for b in range(logits.shape[0]):
        ids = logits[b, :int(seq_len[b])].argmax(-1)
        prev = -1
        seq = []
        for i in ids:
            if i != prev and i != blank:
                seq.append(int(i))
            prev = i
        res.append(seq)
    return res


def clip_by_norm(g, clip):
    """tf.clip_by_norm: g * clip / max(||g||, clip)."""
    n = np.sqrt((g.astype(np.float64) ** 2).sum())
    return (g * (clip / max(n, clip))).astype(g.dtype)


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update (epsilon-hat form): lr_t = lr*sqrt(1-b2^t)/(1-b1^t)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    p = p - lr_t * m / (np.sqrt(v) + eps)
    return p, m, v


def edit_distance_merge(hyp, truth, eos_id):
    """utils/metrics_utils.py:8-41 semantics on python lists: trim at first EOS, merge
    consecutive repeats in BOTH sequences, normalised Levenshtein distance."""
    def prep(s):
        out = []
        for t in s:
            if t == eos_id:
                break
            if not out or out[-1] != t:
                out.append(t)
        return out
    h, t = prep(hyp), prep(truth)
    d = np.arange(len(t) + 1)
    for i in range(1, len(h) + 1):
        prev, d[0] = d[0], i
        for j in range(1, len(t) + 1):
            cur = min(d[j] + 1, d[j - 1] + 1, prev + (h[i - 1] != t[j - 1]))
            prev, d[j] = d[j], cur
    return d[len(t)] / max(len(t), 1)


def ctc_greedy_decoder(logits, lengths):
    """tf.nn.ctc_greedy_decoder(merge_repeated=True) + sparse.to_dense (model_helper.py:351-353): per-frame argmax over the
    V + 1 classes, repeats merged, then the blank -- the LAST class for this op -- removed; rows padded with 0."""
    blank = logits.shape[-1] - 1
    rows = []
    for b in range(logits.shape[0]):
        path = logits[b, :int(lengths[b])].argmax(-1)
        keep = np.ones(len(path), bool)
        keep[1:] = path[1:] != path[:-1]
        seq = path[keep]
        rows.append(seq[seq != blank])
    width = max((len(r) for r in rows), default=0)
    out = np.zeros((len(rows), width), np.int32)
    for b, r in enumerate(rows):
        out[b, :len(r)] = r
    return out
