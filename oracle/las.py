"""Oracle: Listen-Attend-Spell forward pass (numpy restatement; test infrastructure only).

Follows the reference graph code and restates the TensorFlow 1.15.2 ops it calls
(tensorflow is not vendored/installable here -> parity UNPINNED, see oracle/__init__.py):

* las/ops.py:10-20    lstm_cell        -> tf.nn.rnn_cell.LSTMCell  (gate order i,j,f,o; forget_bias=1)
* las/ops.py:23-46    bilstm           -> tf.nn.bidirectional_dynamic_rnn / dynamic_rnn
* las/ops.py:49-65    pyramidal_stack
* las/ops.py:68-87    pyramidal_bilstm
* las/model.py:104-142 listener
* las/model.py:145-202 attend          -> tf.contrib.seq2seq.{Luong,Bahdanau,LuongMonotonic}Attention,
                                          AttentionWrapper
* las/model.py:205-349 speller         -> BasicDecoder + GreedyEmbeddingHelper / TrainingHelper +
                                          dynamic_decode
* utils/training_helper.py:122-153 DenseBinfDecoder (the output projection)

Precision contract.  ``precision='fp32'`` computes everything in float32 like the reference.
``precision='bf16'`` emulates the CUDA path's storage points (DESIGN.md "precision contract"):
weights, layer inputs, stored gate pre-activations, recurrent h, attention keys/values and the
attention vector fed back to the decoder cell are rounded to bfloat16 (round-to-nearest-even);
accumulation, gates, cell state, softmax, the context entering the output projection and the
logits stay float32.  Where the folded-context tensor-core decoder runs (``folded_context_decoder``: bf16,
default wiring, decoder_units and encoder depth multiples of 64) the fed-back context is never formed:
cell 0 adds ``alignments @ VW`` with ``VW = bf16(values @ W0[V:V+D])`` -- the rounding point moves from
the context to VW.
"""
import numpy as np

F32 = np.float32


def round_bf16(x):
    """float32 -> nearest-even bfloat16 -> float32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    rounded = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
    out = rounded.view(np.float32)
    return np.where(np.isfinite(x), out, x).astype(np.float32)


def _q(precision):
    if precision == "fp32":
        return lambda a: np.asarray(a, dtype=F32)
    if precision == "bf16":
        return round_bf16
    raise ValueError(precision)


def folded_context_decoder(hp, B, D, precision):
    """Shapes on which the CUDA path runs decoder_fold.cu (speller.SpellerWeights.tc + dec_fold_plan): its bf16 storage
    points differ from the other decoders' (VW instead of the fed-back context), so the emulation must know."""
    Ud = hp["decoder_units"]
    return (precision == "bf16" and not hp.get("bottom_only") and not hp.get("attention_layer_size")
            and not hp.get("binf_projection") and hp["attention_type"] in ("luong", "bahdanau", "luong_monotonic", "custom")
            and D % 64 == 0 and Ud % 64 == 0 and D <= 2048 and Ud // 4 <= 148 and B <= 128)


def sigmoid(x):
    x = np.asarray(x, dtype=F32)
    return (F32(1) / (F32(1) + np.exp(-x))).astype(F32)


# --------------------------------------------------------------------------------------
# LSTMCell / dynamic_rnn
# --------------------------------------------------------------------------------------
def lstm_cell_step(z, c_prev):
    """Gate math of tf LSTMCell given pre-activations z=[B,4U] (order i,j,f,o; forget_bias 1.0)."""
    U = z.shape[1] // 4
    i, j, f, o = z[:, :U], z[:, U:2 * U], z[:, 2 * U:3 * U], z[:, 3 * U:]
    c = sigmoid(f + F32(1.0)) * c_prev + sigmoid(i) * np.tanh(j).astype(F32)
    h = sigmoid(o) * np.tanh(c).astype(F32)
    return c.astype(F32), h.astype(F32)


def dynamic_rnn(x, lengths, kernel, bias, precision="fp32", reverse=False):
    """tf.nn.dynamic_rnn over one LSTMCell with sequence_length semantics.

    x [B,T,din] f32, lengths [B]; kernel [din+U,4U]; bias [4U].
    Outputs are 0 for t >= len, state is frozen past len; ``reverse`` applies
    reverse_sequence (within each length) on input and output (the bw direction of
    bidirectional_dynamic_rnn).  Returns out [B,T,U], (c,h) final.
    """
    q = _q(precision)
    B, T, din = x.shape
    U = kernel.shape[1] // 4
    wx, wh = q(kernel[:din]), q(kernel[din:])
    xin = q(x)
    # time-parallel input projection (mathematically what TF does per step inside [x,h]@kernel)
    xproj = q(xin.reshape(B * T, din) @ wx + bias.astype(F32)).reshape(B, T, 4 * U)
    c = np.zeros((B, U), F32)
    h = np.zeros((B, U), F32)
    out = np.zeros((B, T, U), F32)
    lengths = np.asarray(lengths)
    for s in range(T):
        active = s < lengths
        if not active.any():
            break
        t_idx = np.where(active, (lengths - 1 - s) if reverse else s, 0)
        z = xproj[np.arange(B), t_idx] + h @ wh
        c_new, h_new = lstm_cell_step(z.astype(F32), c)
        h_new = q(h_new)
        c = np.where(active[:, None], c_new, c)
        h = np.where(active[:, None], h_new, h)
        rows = np.nonzero(active)[0]
        out[rows, t_idx[rows]] = h_new[rows]
    return out, (c, h)


def bilstm(x, lengths, params, scope, precision="fp32", unidirectional=False):
    """las/ops.py:23-46.  ``params`` maps TF variable names to arrays."""
    if unidirectional:
        k = params[scope + "/rnn/lstm_cell/kernel"]
        b = params[scope + "/rnn/lstm_cell/bias"]
        return dynamic_rnn(x, lengths, k, b, precision)
    outs, states = [], []
    for d, rev in (("fw", False), ("bw", True)):
        k = params[f"{scope}/bidirectional_rnn/{d}/lstm_cell/kernel"]
        b = params[f"{scope}/bidirectional_rnn/{d}/lstm_cell/bias"]
        o, st = dynamic_rnn(x, lengths, k, b, precision, reverse=rev)
        outs.append(o)
        states.append(st)
    return tuple(outs), tuple(states)


def pyramidal_stack(outputs, lengths):
    """las/ops.py:49-65."""
    B, T, D = outputs.shape
    if T % 2:
        outputs = np.concatenate([outputs, np.zeros((B, 1, D), outputs.dtype)], axis=1)
    outputs = outputs.reshape(B, -1, 2 * D)
    lengths = np.asarray(lengths)
    return outputs, lengths // 2 + lengths % 2


def pyramidal_bilstm(x, lengths, params, num_layers, precision="fp32", unidirectional=False,
                     scope="listener"):
    """las/ops.py:68-87."""
    outputs = x
    state = None
    for layer in range(num_layers):
        o, state = bilstm(outputs, lengths, params, f"{scope}/bilstm_{layer}", precision, unidirectional)
        outputs = o if unidirectional else np.concatenate(o, -1)
        if layer != 0:
            outputs, lengths = pyramidal_stack(outputs, lengths)
    return (outputs, np.asarray(lengths)), state


def listener(encoder_inputs, source_sequence_length, params, hp, precision="fp32"):
    """las/model.py:104-142 (pyramidal branch + non-pyramidal bidirectional MultiRNNCell branch)."""
    if hp["use_pyramidal"]:
        return pyramidal_bilstm(encoder_inputs, source_sequence_length, params, hp["encoder_layers"],
                                precision, hp.get("unidirectional", False))
    # stacked MultiRNNCell: each direction is an independent L-layer stack over the raw input
    q = _q(precision)
    B, T, _ = encoder_inputs.shape
    lengths = np.asarray(source_sequence_length)
    outs, states = [], []
    dirs = (("fw", False),) if hp.get("unidirectional", False) else (("fw", False), ("bw", True))
    for d, rev in dirs:
        xin = encoder_inputs
        st_d = []
        for l in range(hp["encoder_layers"]):
            base = ("listener/rnn" if hp.get("unidirectional", False)
                    else f"listener/bidirectional_rnn/{d}")
            k = params[f"{base}/multi_rnn_cell/cell_{l}/lstm_cell/kernel"]
            b = params[f"{base}/multi_rnn_cell/cell_{l}/lstm_cell/bias"]
            # NOTE: a MultiRNNCell steps all layers per time step; for a stack of LSTMs with
            # length masking this equals running the layers one after another on the (already
            # direction-reversed) sequence.  We run each layer in the direction's own time order.
            xin, st = dynamic_rnn(xin, lengths, k, b, precision, reverse=rev)
            st_d.append(st)
        outs.append(xin)
        states.append(tuple(st_d))
    out = outs[0] if len(outs) == 1 else np.concatenate(outs, -1)
    # bidirectional_dynamic_rnn returns (fw MultiRNNCell state, bw state); dynamic_rnn the single stack's state
    return (out, lengths), (tuple(states) if len(states) == 2 else states[0])


# --------------------------------------------------------------------------------------
# attention mechanisms (tf.contrib.seq2seq)
# --------------------------------------------------------------------------------------
def _safe_cumprod_exclusive(x):
    tiny = np.finfo(np.float32).tiny
    logx = np.log(np.clip(x, tiny, 1.0)).astype(F32)
    cs = np.cumsum(logx, axis=1, dtype=F32)
    cs = np.concatenate([np.zeros_like(cs[:, :1]), cs[:, :-1]], axis=1)
    return np.exp(cs).astype(F32)


class Attention:
    """values/keys setup of _BaseAttentionMechanism + score/probability fns."""

    def __init__(self, attention_type, memory, memory_len, params, scope, precision="fp32", prefix=None):
        q = self.q = _q(precision)
        self.type = attention_type
        B, Tm, D = memory.shape
        self.mask = (np.arange(Tm)[None, :] < np.asarray(memory_len)[:, None])
        self.values = q(memory * self.mask[:, :, None].astype(F32))
        wm = q(params[f"{scope}/memory_layer/kernel"])
        self.keys = q((self.values.reshape(B * Tm, D) @ wm).reshape(B, Tm, -1))
        pre = prefix or f"{scope}/decoder/attention_wrapper"
        if attention_type in ("bahdanau", "bahdanau_monotonic"):
            self.wq = q(params[f"{pre}/{attention_type}_attention/query_layer/kernel"])
            self.v = params[f"{pre}/{attention_type}_attention/attention_v"].astype(F32)
        elif attention_type == "custom":  # CustomAttention, las/model.py:72-101: relu on the keys and on its own query layer
            self.wq = q(params[f"{pre}/query_layer/kernel"])
            self.keys = np.maximum(self.keys, F32(0))
        elif attention_type not in ("luong", "luong_monotonic"):
            raise NotImplementedError(attention_type)
        if attention_type.endswith("_monotonic"):
            self.score_bias = F32(params[f"{pre}/{attention_type}_attention/attention_score_bias"])

    def initial_alignments(self):
        B, Tm = self.mask.shape
        a = np.zeros((B, Tm), F32)
        if self.type.endswith("_monotonic"):
            a[:, 0] = 1.0
        return a

    def __call__(self, query, prev_alignments):
        """query [B,Ud] (already quantised cell output) -> alignments [B,Tm] f32."""
        if self.type in ("bahdanau", "bahdanau_monotonic"):
            pq = (query @ self.wq).astype(F32)
            score = np.einsum("btu,u->bt", np.tanh(self.keys + pq[:, None, :]).astype(F32), self.v,
                              dtype=F32)
        elif self.type == "custom":
            score = np.einsum("btu,bu->bt", self.keys, np.maximum((query @ self.wq).astype(F32), F32(0)), dtype=F32)
        else:
            score = np.einsum("btu,bu->bt", self.keys, query, dtype=F32)
        if self.type == "bahdanau_monotonic":
            # outside TRAIN the reference asks for mode='hard' (las/model.py:161-164): p = [score > 0] * cumsum(prev),
            # alignments = p * cumprod_exclusive(1 - p)  (tf.contrib.seq2seq.monotonic_attention)
            p = np.where(self.mask & (score + self.score_bias > 0), F32(1), F32(0)) * np.cumsum(prev_alignments, axis=1, dtype=F32)
            one_m = F32(1) - p
            cp = np.concatenate([np.ones_like(p[:, :1]), np.cumprod(one_m, axis=1, dtype=F32)[:, :-1]], axis=1)
            return (p * cp).astype(F32)
        if self.type == "luong_monotonic":
            score = score + self.score_bias
            with np.errstate(over="ignore"):
                p = np.where(self.mask, sigmoid(score), F32(0)).astype(F32)
            cp = _safe_cumprod_exclusive(F32(1) - p)
            return (p * cp * np.cumsum(prev_alignments / np.clip(cp, 1e-10, 1.0), axis=1, dtype=F32)).astype(F32)
        score = np.where(self.mask, score, -np.inf).astype(F32)
        m = score.max(axis=1, keepdims=True)
        e = np.exp(score - m).astype(F32)
        return (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)


# --------------------------------------------------------------------------------------
# AttentionWrapper(MultiRNNCell) + BasicDecoder + helpers + dynamic_decode
# --------------------------------------------------------------------------------------
class Speller:
    """las/model.py:185-200 with attention_layer_size=None: the default wiring (AttentionWrapper around the MultiRNNCell) or
    ``bottom_only`` (GNMT-style AttentionMultiCell, las/model.py:20-69: attention wraps cell 0 only; cell 0's output is the
    NEW attention, every upper cell reads [previous output; OLD attention]; the decoder output is the top cell's h), with
    ``pass_hidden_state`` (las/model.py:259-267: cell l starts from the listener's final state l = fw, bw)."""

    def __init__(self, enc_out, enc_len, params, hp, precision="fp32", scope="speller", encoder_state=None, binf=None,
                 fold_context=None):
        self.q = _q(precision)
        self.hp = hp
        self.scope = scope
        self.att = Attention(hp["attention_type"], enc_out, enc_len, params, scope, precision,
                             prefix=(f"{scope}/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper"
                                     if hp.get("bottom_only") else None))
        self.B, self.Tm, self.D = enc_out.shape
        self.Ud = hp["decoder_units"]
        self.V = hp["target_vocab_size"]
        self.bottom_only = bool(hp.get("bottom_only", False))
        self.init_state = None
        if hp.get("pass_hidden_state") and self.bottom_only:
            # zip(decoder zero_state, encoder_state): cell l <- final (c, h) of the last listener layer, l = 0 fw, 1 bw
            self.init_state = [(np.asarray(c, F32), np.asarray(h, F32)) for (c, h) in encoder_state]
        self.cells = []
        for k in range(hp["decoder_layers"]):
            if self.bottom_only:
                name = (f"{scope}/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper/lstm_cell" if k == 0
                        else f"{scope}/decoder/multi_rnn_cell/cell_{k}/lstm_cell")
            else:
                name = f"{scope}/decoder/attention_wrapper/multi_rnn_cell/cell_{k}/lstm_cell"
            self.cells.append((self.q(params[name + "/kernel"]), params[name + "/bias"].astype(F32)))
        al = (f"{scope}/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper/attention_layer/kernel" if self.bottom_only
              else f"{scope}/decoder/attention_wrapper/attention_layer/kernel")
        self.wal = self.q(params[al]) if hp.get("attention_layer_size") else None
        self.wp = self.q(params[f"{scope}/decoder/projection_layer/kernel"])
        self.bp = params[f"{scope}/decoder/projection_layer/bias"].astype(F32)
        # --binf_projection (las/model.py:240-241,251-257): ``binf`` = binf2phone [n, V]; the decoder is fed the binary-feature
        # column of the previous phone, the 2n-wide attention vector is read as [log p1 | log p0] and mapped to phone scores by
        # transform_binf_to_phones (DenseBinfDecoder with inner_projection_layer=False: its Dense variables are never used)
        self.binf = None if binf is None else np.asarray(binf, F32)
        if hp.get("embedding_size"):
            self.target_embedding = np.asarray(params[f"{scope}/target_embedding"], F32)
        if self.binf is not None:
            assert hp.get("binf_projection") and not self.bottom_only and self.wal is not None  # bottom_only: not restated
            assert self.wal.shape[1] == 2 * self.binf.shape[0], "attention_layer_size must be 2 * binf_count (las/model.py:180-183)"
        self.enc_len = np.asarray(enc_len)
        # folded-context decoder: VW = bf16(values @ W0[V:V+D]) replaces the bf16 copy of the fed-back context
        self.fold = (folded_context_decoder(hp, self.B, self.D, precision) and binf is None) if fold_context is None else bool(fold_context)
        self.nx = int(hp.get("embedding_size") or 0) or self.V  # width of the decoder input (one-hot or target_embedding row)
        if self.fold:
            k0 = self.cells[0][0]
            self.vw = self.q((self.att.values.reshape(self.B * self.Tm, self.D) @ k0[self.nx:self.nx + self.D]).astype(F32)
                             ).reshape(self.B, self.Tm, 4 * self.Ud)

    def zero_state(self):
        cs = [(np.zeros((self.B, self.Ud), F32), np.zeros((self.B, self.Ud), F32)) for _ in self.cells]
        if self.init_state is not None:
            for l, st in enumerate(self.init_state[:len(cs)]):
                cs[l] = st
        st = dict(cells=cs, attention=np.zeros((self.B, self.D if self.wal is None else self.wal.shape[1]), F32),
                  alignments=self.att.initial_alignments())
        if self.fold:
            st["zctx"] = np.zeros((self.B, 4 * self.Ud), F32)
        return st

    def step(self, x, state):
        """AttentionWrapper.call + output projection.  x [B,V] (one-hot or teacher input)."""
        q = self.q
        if self.bottom_only:
            return self._step_bottom_only(x, state)
        inp = np.concatenate([x, state["attention"]], axis=1).astype(F32)
        new_cells = []
        for li, ((k, b), (c, h)) in enumerate(zip(self.cells, state["cells"])):
            if self.fold and li == 0:  # context rows of cell 0 folded into VW: z = x W[:V] + a_{t-1} VW + h W[V+D:] + b
                z = (x.astype(F32) @ k[:self.nx] + state["zctx"] + h @ k[self.nx + self.D:] + b).astype(F32)
            else:
                z = (np.concatenate([inp, h], axis=1) @ k + b).astype(F32)
            c2, h2 = lstm_cell_step(z, c)
            h2 = q(h2)
            new_cells.append((c2, h2))
            inp = h2
        align = self.att(inp, state["alignments"])
        context = np.einsum("bt,btd->bd", align, self.att.values, dtype=F32)
        if self.wal is not None:  # attention_layer_size: attention = Dense([cell_output; context]), no bias
            context = (np.concatenate([inp, context], axis=1) @ self.wal).astype(F32)
        attention = q(context)  # the recurrent feedback copy of the attention vector is rounded ...
        # ... while the projection consumes the f32 context (fp32 mode: q is the identity, so this is
        # exactly DenseBinfDecoder(attention); bf16 mode: one rounding point fewer)
        if self.binf is not None:  # utils/training_helper.py:17-27
            n = self.binf.shape[0]
            logits = (context[:, :n] @ self.binf + context[:, n:2 * n] @ (F32(1) - self.binf)).astype(F32)
        else:
            logits = (context.astype(F32) @ self.wp + self.bp).astype(F32)
        new_state = dict(cells=new_cells, attention=attention, alignments=align)
        if self.fold:
            new_state["zctx"] = np.einsum("bt,btz->bz", align, self.vw, dtype=F32)
        return logits, new_state

    def _step_bottom_only(self, x, state):
        """AttentionMultiCell.__call__ (las/model.py:34-69, use_new_attention=False) + projection of the top cell's output."""
        q = self.q
        old_att = state["attention"]
        (k0, b0), (c, h) = self.cells[0], state["cells"][0]
        z = (np.concatenate([x, old_att, h], axis=1) @ k0 + b0).astype(F32)
        c2, h2 = lstm_cell_step(z, c)
        h2 = q(h2)
        new_cells = [(c2, h2)]
        align = self.att(h2, state["alignments"])
        context = np.einsum("bt,btd->bd", align, self.att.values, dtype=F32)
        if self.wal is not None:  # attention_layer_size: attention = Dense([cell 0 output; context]), no bias
            context = (np.concatenate([h2, context], axis=1) @ self.wal).astype(F32)
        cur = q(context)  # AttentionWrapper(output_attention=True) returns the attention as cell 0's output
        for (k, b), (c, h) in zip(self.cells[1:], state["cells"][1:]):
            z = (np.concatenate([cur, old_att, h], axis=1) @ k + b).astype(F32)
            c2, h2 = lstm_cell_step(z, c)
            h2 = q(h2)
            new_cells.append((c2, h2))
            cur = h2
        out = context.astype(F32) if len(self.cells) == 1 else cur
        logits = (out @ self.wp + self.bp).astype(F32)
        return logits, dict(cells=new_cells, attention=q(context), alignments=align)

    def one_hot(self, ids):
        """embedding_fn (las/model.py:228-246): one-hot ids, or the phone's binary-feature column under --binf_projection."""
        if self.hp.get("embedding_size"):  # las/model.py:230-237
            return self.q(self.target_embedding[ids])
        if self.binf is not None:
            return np.ascontiguousarray(self.binf.T[ids])
        return np.eye(self.V, dtype=F32)[ids]

    def greedy(self):
        """GreedyEmbeddingHelper + dynamic_decode(impute_finished=False) (las/model.py:270-274,337-347)."""
        hp = self.hp
        max_iter = int(np.rint(F32(self.enc_len.max()) * F32(hp.get("decoding_length_factor", 1.0))))
        state = self.zero_state()
        ids = np.full((self.B,), hp["sos_id"], np.int64)
        finished = np.zeros((self.B,), bool) | (0 >= max_iter)
        seq_len = np.zeros((self.B,), np.int32)
        logits_all, ids_all, align_all = [], [], []
        time = 0
        while not finished.all():
            logits, state = self.step(self.one_hot(ids), state)
            ids = logits.argmax(axis=1)
            step_fin = ids == hp["eos_id"]
            next_fin = step_fin | finished
            seq_len = np.where(~finished, time + 1, seq_len).astype(np.int32)
            next_fin |= (time + 1 >= max_iter)
            logits_all.append(logits)
            ids_all.append(ids.astype(np.int32))
            align_all.append(state["alignments"])
            finished = next_fin
            time += 1
        if not logits_all:
            return (np.zeros((self.B, 0, self.V), F32), np.zeros((self.B, 0), np.int32),
                    np.zeros((self.B, 0, self.Tm), F32), seq_len, state)
        return (np.stack(logits_all, 1), np.stack(ids_all, 1), np.stack(align_all, 1), seq_len, state)

    def beam_search(self, beam_width):
        """tf.contrib.seq2seq.BeamSearchDecoder (length_penalty_weight = 0) + dynamic_decode + gather_tree, as
        las/model.py:219-226,298-319 builds it for PREDICT with beam_width > 0 (start tokens = sos; no partial targets).
        ``self`` must have been built on the TILED memory (tile_batch: row b*W + w = utterance b, also for an encoder_state).
        Returns (predicted_ids [B, T, W] after gather_tree, parent_ids [B, T, W], step word ids [B, T, W], final log-probs [B, W],
        final state lengths [B, W], dynamic_decode's sequence lengths [B, W])."""
        hp, W, V = self.hp, int(beam_width), self.V
        assert self.B % W == 0
        B = self.B // W
        eos = hp["eos_id"]
        max_iter = int(np.rint(F32(self.enc_len.max()) * F32(hp.get("decoding_length_factor", 1.0))))
        state = self.zero_state()
        ids = np.full((self.B,), hp["sos_id"], np.int64)
        log_probs = np.full((B, W), -np.inf, F32)
        log_probs[:, 0] = 0.0                               # initialize(): one_hot(0, W, on 0.0, off -inf)
        finished = np.ones((B, W), bool)
        finished[:, 0] = False                              # self._finished = one_hot(0, W, on False, off True)
        lengths = np.zeros((B, W), np.int64)
        seq_len = np.zeros((B, W), np.int32)
        words, parents = [], []
        time = 0
        done = max_iter <= 0
        while not done:
            logits, state = self.step(self.one_hot(ids), state)
            lg = logits.reshape(B, W, V).astype(F32)
            m = lg.max(-1, keepdims=True)
            lsm = (lg - m - np.log(np.exp(lg - m).sum(-1, keepdims=True, dtype=F32))).astype(F32)
            fin_row = np.full((V,), np.finfo(np.float32).min, F32)
            fin_row[eos] = 0.0                              # _mask_probs: a finished beam can only be continued by eos, at no cost
            step_lp = np.where(finished[:, :, None], fin_row[None, None, :], lsm)
            with np.errstate(over="ignore", invalid="ignore"):
                total = (log_probs[:, :, None] + step_lp).astype(F32).reshape(B, W * V)
            order = np.argsort(-total, axis=1, kind="stable")
            idx = order[:, :W]                                           # top_k: ties -> the lower index first
            top = np.take_along_axis(total, order[:, :W + 1], 1).astype(np.float64)
            with np.errstate(invalid="ignore"):
                gaps = top[:, :-1] - top[:, 1:]
            gaps = gaps[np.isfinite(gaps)]
            if gaps.size:                                                # how decisive the selections were (for parity tests)
                self.beam_margin = min(getattr(self, "beam_margin", np.inf), float(gaps.min()))
            new_lp = np.take_along_axis(total, idx, 1)
            word, beam = (idx % V).astype(np.int32), (idx // V).astype(np.int32)
            prev_fin = np.take_along_axis(finished, beam, 1)
            next_fin = prev_fin | (word == eos)
            lengths = np.take_along_axis(lengths, beam, 1) + (~prev_fin).astype(np.int64)
            seq_len = np.where(~finished, time + 1, seq_len).astype(np.int32)   # dynamic_decode, on the slot's previous flag
            flat = (np.arange(B)[:, None] * W + beam).reshape(-1)               # gather the cell state by parent beam
            state = dict({k: v[flat] for k, v in state.items() if k == "zctx"},
                         cells=[(c[flat], h[flat]) for c, h in state["cells"]], attention=state["attention"][flat],
                         alignments=state["alignments"][flat])
            log_probs, finished = new_lp, next_fin
            words.append(word)
            parents.append(beam)
            ids = word.reshape(-1).astype(np.int64)
            time += 1
            done = finished.all() or time >= max_iter
        T = len(words)
        if T == 0:
            z = np.zeros((B, 0, W), np.int32)
            return z, z, z, log_probs, lengths, seq_len
        step_ids, parent_ids = np.stack(words, 1), np.stack(parents, 1)         # [B, T, W]
        out = np.full((B, T, W), eos, np.int32)                                 # gather_tree (beam_search_ops.cc)
        for b in range(B):
            max_len = min(T, int(lengths[b].max()))
            if max_len <= 0:
                continue
            for w in range(W):
                out[b, max_len - 1, w] = step_ids[b, max_len - 1, w]
                parent = parent_ids[b, max_len - 1, w]
                for level in range(max_len - 2, -1, -1):
                    out[b, level, w] = step_ids[b, level, parent]
                    parent = parent_ids[b, level, parent]
                fin = False
                for t in range(max_len):
                    if fin:
                        out[b, t, w] = eos
                    elif out[b, t, w] == eos:
                        fin = True
        return out, parent_ids, step_ids, log_probs, lengths, seq_len

    def teacher_forced(self, targets_inputs, target_len):
        """TrainingHelper + dynamic_decode (las/model.py:276-296 with sampling_probability=0)."""
        target_len = np.asarray(target_len)
        steps = int(target_len.max())
        if self.hp.get("max_symbols", -1) > 0:
            steps = min(steps, self.hp["max_symbols"])
        state = self.zero_state()
        logits_all = []
        for t in range(steps):
            logits, state = self.step(self.one_hot(targets_inputs[:, t]), state)
            logits_all.append(logits)
        return np.stack(logits_all, 1), state


# --------------------------------------------------------------------------------------
# whole-path helpers
# --------------------------------------------------------------------------------------
def predict(features, lengths, params, hp, precision="fp32"):
    """model_helper.py:165-297 PREDICT predictions dict (greedy, beam_width=0)."""
    (enc_out, enc_len), enc_state = listener(features, lengths, params, hp, precision)
    sp = Speller(enc_out, enc_len, params, hp, precision, encoder_state=enc_state)
    logits, ids, align, seq_len, _ = sp.greedy()
    e = np.exp(logits - logits.max(-1, keepdims=True)) if logits.size else logits
    probs = e / e.sum(-1, keepdims=True) if logits.size else logits
    out = dict(encoder_out=enc_out, source_length=enc_len, sample_ids=ids, alignment=align, probs=probs, logits=logits,
               final_sequence_length=seq_len)
    emb = encoder_embedding(enc_state)
    if emb is not None:
        out["embedding"] = emb
    return out


def encoder_embedding(enc_state):
    """model_helper.py:258-268: ``tf.concat([x.c for x in encoder_state])`` works when the state is a sequence of LSTMStateTuples
    (pyramidal bidirectional: (fw, bw); stacked unidirectional: one per layer); a single LSTMStateTuple (pyramidal
    unidirectional) takes the ``encoder_state.c`` branch; the stacked bidirectional state (a tuple of per-direction tuples)
    fails both and the prediction has no 'embedding'."""
    is_pair = lambda s: isinstance(s, (tuple, list)) and len(s) == 2 and all(isinstance(t, np.ndarray) for t in s)
    if is_pair(enc_state):
        emb_c, emb_h = enc_state
    elif isinstance(enc_state, (tuple, list)) and len(enc_state) and all(is_pair(s) for s in enc_state):
        emb_c = np.concatenate([s[0] for s in enc_state], axis=1)
        emb_h = np.concatenate([s[1] for s in enc_state], axis=1)
    else:
        return None
    return np.stack([emb_c, emb_h], 1)
