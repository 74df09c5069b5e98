"""Oracle: the TRAIN graph of las_model_fn as differentiable torch (CPU) code -- test infrastructure only.

The numpy oracle (oracle/las.py, oracle/losses.py) restates forward values; this file restates the same
functions with torch ops so that autograd yields the gradients the CUDA backward kernels are checked against
(float64 by default: a reference for the floating-point backward pass).  It is
cross-checked against the numpy oracle's forward values in tests/test_oracle_train.py.

Follows (reference file:line):
* las/ops.py:10-87             lstm_cell / bilstm / pyramidal_stack / pyramidal_bilstm
* las/model.py:145-202         attend (luong, bahdanau; default wiring)
* las/model.py:205-296,344-349 speller TRAIN branch: TrainingHelper / TrainingSigmoidHelper + BasicDecoder +
                               dynamic_decode, DenseBinfDecoder projection (utils/training_helper.py:122-153)
* model_helper.py:20-30        compute_loss TRAIN, :81-105 compute_loss_sigmoid TRAIN, :347-358 CTC head
* model_helper.py:403-417      L2 regulariser, per-tensor clip_by_norm(2), Adam
Parameters are a dict of TF variable names -> tensors in the TF checkpoint layout.
"""
import torch

GRAD_NORM = 2.0  # model_helper.py:16


def _cell(z, c_prev):
    U = z.shape[1] // 4
    i, j, f, o = z[:, :U], z[:, U:2 * U], z[:, 2 * U:3 * U], z[:, 3 * U:]
    c = torch.sigmoid(f + 1.0) * c_prev + torch.sigmoid(i) * torch.tanh(j)
    h = torch.sigmoid(o) * torch.tanh(c)
    return c, h


def dynamic_rnn(x, lengths, kernel, bias, reverse=False):
    """tf.nn.dynamic_rnn with sequence_length (outputs 0 and state frozen past each length)."""
    B, T, din = x.shape
    U = kernel.shape[1] // 4
    xproj = x.reshape(B * T, din) @ kernel[:din] + bias
    xproj = xproj.reshape(B, T, 4 * U)
    wh = kernel[din:]
    c = x.new_zeros((B, U))
    h = x.new_zeros((B, U))
    ar = torch.arange(B)
    steps = int(lengths.max())
    rows_t = []
    for s in range(steps):
        active = (s < lengths)
        t_idx = torch.where(active, (lengths - 1 - s) if reverse else torch.full_like(lengths, s),
                            torch.zeros_like(lengths))
        z = xproj[ar, t_idx] + h @ wh
        c_new, h_new = _cell(z, c)
        m = active[:, None].to(x.dtype)
        c = m * c_new + (1 - m) * c
        h = m * h_new + (1 - m) * h
        rows_t.append((t_idx, active, h_new))
    # scatter without in-place autograd trouble: build per-time outputs
    pieces = x.new_zeros((B, T, U))
    idx_b, idx_t, vals = [], [], []
    for t_idx, active, h_new in rows_t:
        rows = torch.nonzero(active)[:, 0]
        idx_b.append(rows)
        idx_t.append(t_idx[rows])
        vals.append(h_new[rows])
    if idx_b:
        ib, it, vv = torch.cat(idx_b), torch.cat(idx_t), torch.cat(vals)
        pieces = pieces.index_put((ib, it), vv)
    return pieces, (c, h)


def pyramidal_bilstm(x, lengths, params, num_layers, scope="listener", masks=None):
    """``masks[(layer, dir)]`` [B,T,din]: input-dropout multipliers (1/keep or 0) of each cell's DropoutWrapper
    (las/ops.py:14-18), supplied by the caller so that the stochastic op is checked on identical masks."""
    outputs = x
    for layer in range(num_layers):
        outs = []
        for di, (d, rev) in enumerate((("fw", False), ("bw", True))):
            k = params[f"{scope}/bilstm_{layer}/bidirectional_rnn/{d}/lstm_cell/kernel"]
            b = params[f"{scope}/bilstm_{layer}/bidirectional_rnn/{d}/lstm_cell/bias"]
            xin = outputs if masks is None else outputs * masks[(layer, di)]
            o, _ = dynamic_rnn(xin, lengths, k, b, reverse=rev)
            outs.append(o)
        outputs = torch.cat(outs, -1)
        if layer != 0:
            B, T, D = outputs.shape
            if T % 2:
                outputs = torch.cat([outputs, outputs.new_zeros((B, 1, D))], 1)
            outputs = outputs.reshape(B, -1, 2 * D)
            lengths = lengths // 2 + lengths % 2
    return outputs, lengths


def listener(x, lengths, params, hp, masks=None, return_state=False):
    """las/model.py:104-142: pyramidal (las/ops.py:68-87) or stacked MultiRNNCell listener, bi- or unidirectional.
    ``masks[(layer, dir)]``: input-dropout multipliers over the WHOLE layer input; a stacked cell (layer >= 1) reads the
    column slice of its own direction."""
    uni = bool(hp.get("unidirectional", False))
    L, U = hp["encoder_layers"], hp["encoder_units"]
    dirs = (("rnn", False),) if uni else (("bidirectional_rnn/fw", False), ("bidirectional_rnn/bw", True))
    outputs = x
    states = []
    for layer in range(L):
        outs, states = [], []
        for di, (d, rev) in enumerate(dirs):
            if hp["use_pyramidal"]:
                name = f"listener/bilstm_{layer}/{d}/lstm_cell"
            else:
                name = f"listener/{d}/multi_rnn_cell/cell_{layer}/lstm_cell"
            xin = outputs if masks is None else outputs * masks[(layer, di)]
            if not hp["use_pyramidal"] and layer > 0:
                xin = xin[..., di * U:(di + 1) * U]
            o, st = dynamic_rnn(xin, lengths, params[name + "/kernel"], params[name + "/bias"], reverse=rev)
            outs.append(o)
            states.append(st)
        outputs = torch.cat(outs, -1)
        if hp["use_pyramidal"] and layer != 0:
            B, T, D = outputs.shape
            if T % 2:
                outputs = torch.cat([outputs, outputs.new_zeros((B, 1, D))], 1)
            outputs = outputs.reshape(B, -1, 2 * D)
            lengths = lengths // 2 + lengths % 2
    if return_state:  # final (c, h) of the last layer, one per direction (encoder_state of las/ops.py:68-87)
        return outputs, lengths, tuple(states)
    return outputs, lengths


def monotonic_attention(p_choose, previous):
    """tf.contrib.seq2seq.monotonic_attention(mode='parallel') as LuongMonotonicAttention uses it (las/model.py:157-158):
    a = p * cumprod_excl(1 - p) * cumsum(previous / clip(cumprod_excl(1 - p), 1e-10, 1)), the cumprod in log space with the
    argument clipped to [tiny, 1] (safe_cumprod)."""
    tiny = torch.finfo(p_choose.dtype).tiny
    logs = torch.log(torch.clamp(1 - p_choose, tiny, 1.0))
    cp = torch.exp(torch.cumsum(logs, 1) - logs)  # exclusive
    return p_choose * cp * torch.cumsum(previous / torch.clamp(cp, 1e-10, 1.0), 1)


def speller_train(enc_out, enc_len, dec_inputs, params, hp, scope="speller", masks=None, encoder_state=None, sampling=None,
                  fed_inputs=None, score_noise=None, binf=None, attention_out=None, table=None):
    """Teacher-forced decode.  dec_inputs [B,L,E] float (one-hot ids, or binary-feature vectors for the
    binary_outputs speller).  Returns logits [B,L,n_out] (n_out = projection kernel columns).
    ``masks``: input-dropout multipliers of the decoder cells: 'x' [B,L,E] and 'att' [B,L,D] (slot t multiplies
    attention_{t-1}) for cell 0's input [x_t; attention_{t-1}], ('h', l) [B,L,Ud] for the output of layer l feeding l+1.
    ``score_noise`` [B,L,Tm]: bahdanau_monotonic in TRAIN mode adds sigmoid_noise * N(0,1) to the scores (las/model.py:161-162);
    the caller passes the (already scaled) deviates -- or a callable (B, L, Tm) -> deviates -- so that the stochastic op is
    replayed exactly (None = no noise).
    ``binf`` [n, V] (--binf_projection, las/model.py:251-257): the projection is transform_binf_to_phones of the 2n-wide attention
    vector instead of the Dense layer; the attention vectors [B,L,2n] are appended to the list ``attention_out`` (they feed
    compute_log_probs_loss, model_helper.py:243-245,330-331)."""
    B, Tm, D = enc_out.shape
    if callable(score_noise):
        score_noise = score_noise(B, dec_inputs.shape[1], Tm)
    Ud = hp["decoder_units"]
    att_type = hp["attention_type"]
    mask = (torch.arange(Tm)[None, :] < enc_len[:, None])
    values = enc_out * mask[:, :, None].to(enc_out.dtype)
    keys = values @ params[f"{scope}/memory_layer/kernel"]
    bottom = bool(hp.get("bottom_only", False))
    if bottom:  # AttentionMultiCell (las/model.py:20-69)
        pre = f"{scope}/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper"
        names = [f"{pre}/lstm_cell"] + [f"{scope}/decoder/multi_rnn_cell/cell_{k}/lstm_cell" for k in range(1, hp["decoder_layers"])]
        cells = [(params[n + "/kernel"], params[n + "/bias"]) for n in names]
    else:
        pre = f"{scope}/decoder/attention_wrapper"
        cells = [(params[f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell/kernel"],
                  params[f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell/bias"]) for k in range(hp["decoder_layers"])]
    wp = params[f"{scope}/decoder/projection_layer/kernel"]
    bp = params[f"{scope}/decoder/projection_layer/bias"]
    state = [(enc_out.new_zeros((B, Ud)), enc_out.new_zeros((B, Ud))) for _ in cells]
    if bottom and hp.get("pass_hidden_state") and encoder_state is not None:  # las/model.py:259-267
        for l, st in enumerate(encoder_state[:len(state)]):
            state[l] = st
    w_al = params[f"{pre}/attention_layer/kernel"] if hp.get("attention_layer_size") else None
    attention = enc_out.new_zeros((B, D if w_al is None else w_al.shape[1]))
    logits = []
    neg_inf = torch.tensor(float("-inf"), dtype=enc_out.dtype)
    mono = att_type.endswith("_monotonic")
    if mono:  # the alignments are a recurrent state, initialised to a dirac at frame 0 (_BaseMonotonicAttentionMechanism)
        align = torch.zeros((B, Tm), dtype=enc_out.dtype)
        align[:, 0] = 1.0
        score_bias = params[f"{pre}/{att_type}_attention/attention_score_bias"]
    if att_type == "custom":  # CustomAttention (las/model.py:72-101)
        keys = torch.relu(keys)

    def attend(query, t, align):
        if att_type in ("bahdanau", "bahdanau_monotonic"):
            pq = query @ params[f"{pre}/{att_type}_attention/query_layer/kernel"]
            score = (torch.tanh(keys + pq[:, None, :]) * params[f"{pre}/{att_type}_attention/attention_v"]).sum(-1)
        elif att_type == "custom":
            score = torch.einsum("btu,bu->bt", keys, torch.relu(query @ params[f"{pre}/query_layer/kernel"]))
        elif att_type in ("luong", "luong_monotonic"):
            score = torch.einsum("btu,bu->bt", keys, query)
        else:
            raise NotImplementedError(att_type)
        if not mono:
            return torch.softmax(torch.where(mask, score, neg_inf), dim=1)
        score = score + score_bias
        if score_noise is not None:
            score = score + torch.as_tensor(score_noise[:, t], dtype=score.dtype)
        return monotonic_attention(torch.where(mask, torch.sigmoid(score), torch.zeros_like(score)), align)
    # scheduled sampling (ScheduledEmbeddingTrainingHelper, las/model.py:279-288): ``sampling`` = (selected [B,S] bool, gumbel
    # [B,S,V]); where selected[b,t], the input of step t+1 becomes one_hot(argmax(logits_t + gumbel_t)) -- a draw from
    # Categorical(logits_t) by Gumbel-max, with the caller's noise so that the stochastic op is replayed exactly.  The inputs
    # that were actually fed are appended to ``fed_inputs`` (a list) when given.
    dec_inputs = [dec_inputs[:, t] for t in range(dec_inputs.shape[1])]

    def maybe_sample(t, logits_t):
        if sampling is None or t + 1 >= len(dec_inputs):
            return
        sel = torch.as_tensor(sampling[0][:, t])
        ids = (logits_t.detach() + torch.as_tensor(sampling[1][:, t], dtype=logits_t.dtype)).argmax(-1)
        if table is not None:  # embedding_fn other than one-hot: the sampled id feeds its embedding row (differentiable wrt the table)
            drawn = table[ids]
        else:
            drawn = torch.nn.functional.one_hot(ids, logits_t.shape[-1]).to(logits_t.dtype)
        dec_inputs[t + 1] = torch.where(sel[:, None], drawn, dec_inputs[t + 1])

    for t in range(len(dec_inputs)):
        if fed_inputs is not None:
            fed_inputs.append(dec_inputs[t])
        if bottom:
            old = attention
            (k0, b0), (c, h) = cells[0], state[0]
            if masks is None:
                inp0 = torch.cat([dec_inputs[t], old, h], 1)
            else:  # DropoutWrapper on every cell: cell 0 drops [x_t; attention_{t-1}], cell l >= 1 its [output below; old attention]
                inp0 = torch.cat([dec_inputs[t] * masks["x"][:, t], old * masks["att"][:, t], h], 1)
            c0, h0 = _cell(inp0 @ k0 + b0, c)
            new_state = [(c0, h0)]
            align = attend(h0, t, align if mono else None)
            attention = torch.einsum("bt,btd->bd", align, values)
            if w_al is not None:  # attention_layer_size: attention = Dense([cell 0 output; context]), no bias
                attention = torch.cat([h0, attention], 1) @ w_al
            cur = attention
            for li, ((k, b), (c, h)) in enumerate(zip(cells[1:], state[1:]), start=1):
                xin = torch.cat([cur, old], 1)
                if masks is not None:
                    xin = xin * masks[("in", li)][:, t]
                c2, h2 = _cell(torch.cat([xin, h], 1) @ k + b, c)
                new_state.append((c2, h2))
                cur = h2
            state = new_state
            logits.append(cur @ wp + bp)
            maybe_sample(t, logits[-1])
            continue
        if masks is None:
            inp = torch.cat([dec_inputs[t], attention], 1)
        else:
            inp = torch.cat([dec_inputs[t] * masks["x"][:, t], attention * masks["att"][:, t]], 1)
        new_state = []
        for li, ((k, b), (c, h)) in enumerate(zip(cells, state)):
            z = torch.cat([inp, h], 1) @ k + b
            c2, h2 = _cell(z, c)
            new_state.append((c2, h2))
            inp = h2
            if masks is not None and li + 1 < len(cells):
                inp = h2 * masks[("h", li)][:, t]
        state = new_state
        align = attend(inp, t, align if mono else None)
        attention = torch.einsum("bt,btd->bd", align, values)
        if w_al is not None:  # attention_layer_size: attention = Dense([cell output; context]), no bias
            attention = torch.cat([inp, attention], 1) @ w_al
        if attention_out is not None:
            attention_out.append(attention)
        if binf is not None:  # utils/training_helper.py:17-27
            nb = binf.shape[0]
            logits.append(attention[:, :nb] @ binf + attention[:, nb:2 * nb] @ (1 - binf))
        else:
            logits.append(attention @ wp + bp)
        maybe_sample(t, logits[-1])
    return torch.stack(logits, 1)


def sequence_loss(logits, targets, weights):
    lp = torch.log_softmax(logits, -1)
    ce = -lp.gather(-1, targets[..., None].long())[..., 0]
    return (ce * weights).sum() / (weights.sum() + 1e-12)


def sequence_loss_sigmoid(logits, targets, weights):
    x, z = logits, targets
    ce = (torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs()))).mean(-1)
    return (ce * weights).sum() / (weights.sum() + 1e-12)


def compute_log_probs_loss(outputs):
    """model_helper.py:132-146 (the normalisation constant carries no gradient: tf.stop_gradient)."""
    n = outputs.shape[-1] // 2
    l1, l0 = outputs[..., :n], outputs[..., n:2 * n]
    c = (-(l1 + l0) / 2).detach()
    loss = ((torch.exp(l1 + c) + torch.exp(l0 + c)) / torch.exp(c) - 1).abs() + torch.relu(l1) + torch.relu(l0)
    return loss.mean()


def ctc_loss(logits, labels, label_length, logit_length, blank=0):
    """Same recursion as oracle/losses.py:ctc_loss, differentiable.  Returns [B]."""
    out = []
    for b in range(logits.shape[0]):
        T, L = int(logit_length[b]), int(label_length[b])
        lp = torch.log_softmax(logits[b, :T], -1)
        ext = torch.full((2 * L + 1,), blank, dtype=torch.long)
        ext[1::2] = labels[b, :L].long()
        S = 2 * L + 1
        ninf = lp.new_full((1,), -1e30)  # finite stand-in for -inf: exp() underflows to 0, autograd stays NaN-free
        can_skip = torch.zeros((S,), dtype=torch.bool)
        can_skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
        alpha = torch.cat([lp[0, ext[:2]], ninf.expand(max(S - 2, 0))])[:S]
        for t in range(1, T):
            a1 = torch.cat([ninf, alpha[:-1]])
            a2 = torch.cat([ninf, ninf, alpha[:-2]])[:S]
            a2 = torch.where(can_skip, a2, ninf.expand(S))
            alpha = torch.logsumexp(torch.stack([alpha, a1, a2]), 0) + lp[t, ext]
        tail = alpha[-1] if S == 1 else torch.logsumexp(alpha[-2:], 0)
        out.append(-tail)
    return torch.stack(out)


def train_loss(params, features, lengths, labels, hp, binf=None, masks=None, sampling=None, score_noise=None, sampling_binf=None):
    """las_model_fn(mode=TRAIN) loss (model_helper.py:165-358, 411-413) with dropout = 0 and
    sampling_probability = 0.  ``binf`` [n, V] enables the multitask binary-feature speller.
    Returns (total loss incl. L2, dict of the parts)."""
    dt = features.dtype
    masks = masks or {}
    enc_out, enc_len, enc_state = listener(features, lengths, params, hp, masks=masks.get("listener"), return_state=True)
    tin, tout, tlen = labels["targets_inputs"], labels["targets_outputs"], labels["target_sequence_length"]
    L = tin.shape[1]
    w = (torch.arange(L)[None, :] < tlen[:, None]).to(dt)
    parts = {}
    loss = 0.0
    V = hp["target_vocab_size"]
    if not hp.get("binary_outputs") or hp.get("multitask"):
        if hp.get("embedding_size"):  # las/model.py:230-237
            dec_in = params["speller/target_embedding"][tin.long()]
        else:
            dec_in = torch.nn.functional.one_hot(tin.long(), V).to(dt)
        logits = speller_train(enc_out, enc_len, dec_in, params, hp, table=params["speller/target_embedding"] if hp.get("embedding_size") else None,
                               masks=masks.get("speller"), encoder_state=enc_state, sampling=sampling, score_noise=score_noise)
        parts["ce"] = sequence_loss(logits, tout, w)
        parts["logits"] = logits
        loss = loss + parts["ce"]
    if hp.get("binary_outputs") and hp.get("binf_projection"):
        # model_helper.py:221-227 (phone ids in, embedded as binary-feature columns), :243-245 (raw attention split off),
        # :326-331 (softmax CE on the transformed logits + the log-probability regulariser)
        # --binf_trainable: the map is the variable 'binf2phone' (model_helper.py:181-186), read by the embedding and the projection
        M = params["binf2phone"] if hp.get("binf_trainable") else torch.as_tensor(binf, dtype=dt)
        atts = []
        logits_b = speller_train(enc_out, enc_len, M.t()[tin.long()], params, hp, scope="speller_binf",
                                 masks=masks.get("speller_binf"), encoder_state=enc_state, binf=M, attention_out=atts,
                                 sampling=sampling_binf, table=M.t())
        parts["ce_binf"] = sequence_loss(logits_b, tout, w)
        parts["log_probs_reg"] = compute_log_probs_loss(torch.stack(atts, 1))
        parts["logits_binf"] = logits_b
        loss = loss + parts["ce_binf"] + parts["log_probs_reg"] * hp.get("binf_projection_reg_weight", 1.0)
    elif hp.get("binary_outputs"):
        bt = torch.as_tensor(binf, dtype=dt).t()  # [V, n]
        logits_b = speller_train(enc_out, enc_len, bt[tin.long()], params, hp, scope="speller_binf",
                                 masks=masks.get("speller_binf"), encoder_state=enc_state)
        parts["ce_binf"] = sequence_loss_sigmoid(logits_b, bt[tout.long()], w)
        parts["logits_binf"] = logits_b
        loss = loss + parts["ce_binf"]
    if hp.get("ctc_weight", -1.0) > 0:
        cl = enc_out @ params["ctc_logits/kernel"] + params["ctc_logits/bias"]
        parts["ctc"] = ctc_loss(cl, tout, tlen, enc_len).mean()
        loss = loss + parts["ctc"] * hp["ctc_weight"]
    parts["audio_loss"] = loss
    reg = sum((p * p).sum() for p in params.values()) * (0.5 * hp.get("l2_reg_scale", 0.0))
    parts["encoder_out"] = enc_out
    return loss + reg, parts


def clip_and_adam(params, grads, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """model_helper.py:416-417: per-tensor tf.clip_by_norm(g, 2) then tf.train.AdamOptimizer (epsilon-hat form)."""
    out_p, out_m, out_v = {}, {}, {}
    lr_t = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
    for k, p in params.items():
        g = grads[k]
        n = g.pow(2).sum().sqrt()
        g = g * (GRAD_NORM / torch.clamp(n, min=GRAD_NORM))
        out_m[k] = b1 * m[k] + (1 - b1) * g
        out_v[k] = b2 * v[k] + (1 - b2) * g * g
        out_p[k] = p - lr_t * out_m[k] / (out_v[k].sqrt() + eps)
    return out_p, out_m, out_v
