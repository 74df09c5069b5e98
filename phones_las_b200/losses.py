"""Loss heads of the model function on the GPU (forward values).

Mirrors ``compute_loss`` (model_helper.py:20-78), ``sequence_loss_sigmoid`` / ``compute_loss_sigmoid``
(model_helper.py:81-130, TRAIN branch) and the CTC head (model_helper.py:347-358: Dense(V+1) on the encoder
outputs, ``tf.nn.ctc_loss_v2`` with blank index 0, mean over the batch).  The padding / masking glue is torch on
the device; the arithmetic runs in csrc/losses.cu through the C-ABI.  Backward passes are not built yet.
"""
import torch

from . import _lib


def sequence_mask(lengths, maxlen):
    ar = torch.arange(maxlen, device=lengths.device)
    return (ar[None, :] < lengths[:, None].to(ar.dtype)).to(torch.float32)


def _weighted_ce(fn, a, b_, weights, n_tok, inner):
    L = _lib.lib()
    ce = torch.empty((n_tok,), dtype=torch.float32, device=a.device)
    out3 = torch.empty((3,), dtype=torch.float32, device=a.device)
    with _lib.stage("loss"):
        _lib.check(fn(_lib.ptr(a), _lib.ptr(b_), _lib.ptr(weights) if weights is not None else None, n_tok, inner,
                      _lib.ptr(ce), _lib.ptr(out3), _lib.stream_ptr()))
    _lib.count_launches(2)
    return out3[0], ce


def sequence_loss(logits, targets, weights):
    """tf.contrib.seq2seq.sequence_loss defaults: sum(CE*w) / (sum(w) + 1e-12) over batch x time."""
    B, T, V = logits.shape
    lg = logits.to(torch.float32).contiguous()
    tg = targets.to(torch.int32).contiguous()
    w = weights.to(torch.float32).contiguous()
    loss, ce = _weighted_ce(_lib.lib().plas_seq_ce_fwd, lg, tg, w, B * T, V)
    return loss, ce.view(B, T)


def compute_loss(logits, targets, final_sequence_length, target_sequence_length, mode, eos_id):
    """model_helper.py:20-78."""
    if mode == "train":
        T = targets.shape[1]
        if logits.shape[1] < T:
            logits = torch.nn.functional.pad(logits, (0, 0, 0, T - logits.shape[1]))
        w = sequence_mask(target_sequence_length, T)
        return sequence_loss(logits[:, :T], targets, w)[0]
    max_ts = int(target_sequence_length.max().item())
    max_fs = int(final_sequence_length.max().item())
    L = max(max_ts, max_fs)
    logits = logits[:, :max_fs]
    if targets.shape[1] < L:
        targets = torch.nn.functional.pad(targets, (0, L - targets.shape[1]), value=eos_id)
    if logits.shape[1] < L:
        logits = torch.nn.functional.pad(logits, (0, 0, 0, L - logits.shape[1]))
    seq_len = torch.maximum(target_sequence_length, final_sequence_length)
    w = sequence_mask(seq_len, L)
    return sequence_loss(logits[:, :L], targets[:, :L], w)[0]


def sequence_loss_sigmoid(logits, targets, weights):
    """model_helper.py:81-95."""
    B, T, n = logits.shape
    lg = logits.to(torch.float32).contiguous()
    tg = targets.to(torch.float32).contiguous()
    w = weights.to(torch.float32).contiguous()
    return _weighted_ce(_lib.lib().plas_sigmoid_ce_fwd, lg, tg, w, B * T, n)[0]


def compute_loss_sigmoid_train(logits, targets_binf, target_sequence_length):
    """model_helper.py:98-105 (TRAIN branch)."""
    return sequence_loss_sigmoid(logits, targets_binf, sequence_mask(target_sequence_length, logits.shape[1]))


def ctc_loss(logits, labels, label_length, logit_length, blank=0):
    """tf.nn.ctc_loss_v2 dense-label path -> per-utterance negative log-likelihood [B]."""
    L = _lib.lib()
    B, T, Cn = logits.shape
    lg = logits.to(torch.float32).contiguous()
    lab = labels.to(torch.int32).contiguous()
    ll = label_length.to(torch.int32).contiguous()
    tl = logit_length.to(torch.int32).contiguous()
    out = torch.empty((B,), dtype=torch.float32, device=lg.device)
    with _lib.stage("loss"):
        _lib.check(L.plas_ctc_fwd(_lib.ptr(lg), _lib.ptr(lab), _lib.ptr(ll), _lib.ptr(tl), B, T, Cn, lab.shape[1], blank,
                                  _lib.ptr(out), _lib.stream_ptr()))
    _lib.count_launches(1)
    return out


def ctc_head(encoder_out, source_length, targets, target_sequence_length, kernel, bias):
    """model_helper.py:347-358: Dense(V+1) ('ctc_logits/{kernel,bias}') on the encoder outputs, CTC with blank 0,
    reduce_mean over the batch.  ``kernel`` [D, V+1], ``bias`` [V+1] float32 device tensors."""
    B, Tm, D = encoder_out.shape
    x = encoder_out.to(torch.float32).reshape(B * Tm, D).contiguous()
    wt = kernel.t().contiguous()
    logits = torch.empty((B * Tm, wt.shape[0]), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with _lib.stage("loss"):
        _lib.check(L.plas_gemm_f32(_lib.ptr(x), B * Tm, D, D, _lib.ptr(wt), wt.shape[0], D, _lib.ptr(bias), _lib.ptr(logits),
                                   wt.shape[0], _lib.stream_ptr()))
    _lib.count_launches(1)
    logits = logits.view(B, Tm, -1)
    per_utt = ctc_loss(logits, targets, target_sequence_length, source_length, blank=0)
    return per_utt.mean(), logits
