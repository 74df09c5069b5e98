"""Speller operator: attention decoder (greedy / teacher-forced / beam search) on the GPU.

Mirrors ``las.model.speller`` (reference las/model.py:205-349):
``speller(encoder_outputs, encoder_state, decoder_inputs, source_sequence_length,
target_sequence_length, mode, hparams)`` -> ``(BasicDecoderOutput(rnn_output, sample_id),
final_state, final_sequence_length)`` (``FinalBeamSearchDecoderOutput`` with ``beam_width > 0``).  The attention memory is
prepared once per batch (length masking + memory_layer GEMM, las/model.py:168-169).  Two decoder families sit underneath:

* the fused decoders (``plas_decoder_fwd``: the whole decode loop inside one persistent kernel; bf16 tensor-core or SIMT) for the
  default wiring with luong / bahdanau / luong_monotonic attention -- the BASELINE configurations -- and, on the folded bf16
  tensor-core kernel, custom attention;
* the fp32 step-kernel decoder (``plas_decoder_infer_f32``, a host loop over the training path's step kernels on the TF weight
  layout) for reference precision and for every other variant: bahdanau_monotonic (mode 'hard'), custom attention (fp32), the
  AttentionMultiCell wiring (bottom_only, pass_hidden_state), attention_layer_size, --binf_projection, beam search.

``embedding_size`` folds into the weights of both (``fold_embedding``).  DESIGN.md section 8 lists what is not built.
"""
import ctypes as C
import os
from collections import namedtuple

import numpy as np
import torch

from . import _lib, packing

BasicDecoderOutput = namedtuple("BasicDecoderOutput", ["rnn_output", "sample_id"])
SpellerState = namedtuple("SpellerState", ["alignment_history", "n_steps"])


def _round_up(x, m):
    return (x + m - 1) // m * m


def fold_embedding(params, hp, scope, kern0):
    """embedding_size != 0 (las/model.py:230-237): the decoder input is a row of ``<scope>/target_embedding`` [V, E] instead of a
    one-hot.  Cell 0's first E kernel rows only ever see such rows, so the lookup folds into the weights: the V 'one-hot' rows
    the decoders read become target_embedding @ kernel[:E]."""
    E = int(hp.get("embedding_size") or 0)
    if not E:
        return kern0
    emb = np.asarray(params[f"{scope}/target_embedding"], np.float64)
    assert emb.shape == (hp["target_vocab_size"], E), emb.shape
    return np.concatenate([(emb @ kern0[:E].astype(np.float64)).astype(np.float32), kern0[E:]], 0)


class SpellerWeights:
    """Device-resident, kernel-layout copy of the ``speller/`` variables."""

    def __init__(self, params, hp, enc_depth, precision="fp32", device="cuda", scope="speller", binf=None):
        """``binf`` (binf2phone [n, V], numpy): the --binf_projection wiring of the 'speller_binf' scope (las/model.py:240-241,
        251-257): inputs are the previous phone's binary-feature column, the 2n-wide attention vector is mapped to phone scores
        by transform_binf_to_phones.  Both are linear, so they fold into the weights the step kernels already read: the V
        embedding rows of cell 0's kernel become M^T W_0[:n], the projection becomes the constant [M; 1 - M] with a zero bias."""
        _lib.require_cuda()
        self._params, self._device, self._train_state, self.scope = params, device, None, scope  # TRAIN mode: train_state()
        self.binf_projection = binf is not None
        if self.binf_projection:
            if not hp.get("binf_projection") or hp.get("bottom_only"):
                raise NotImplementedError("a binf2phone matrix needs --binf_projection and the default decoder wiring")
            self._init_attention_layer(params, hp, enc_depth, precision, device, scope, binf=np.asarray(binf, np.float32))
            return
        self.bottom_only = bool(hp.get("bottom_only"))
        self.pass_hidden_state = bool(hp.get("pass_hidden_state")) and self.bottom_only  # las/model.py:260 needs both
        if self.bottom_only:
            self._init_bottom_only(params, hp, enc_depth, precision, device, scope)
            return
        if hp["attention_type"] not in _lib.ATT_CODES:
            raise NotImplementedError(f"attention_type={hp['attention_type']}")
        # custom attention also runs on the folded bf16 tensor-core decoder (decoder_fold.cu) when that kernel's shape rules hold
        custom_fused = (hp["attention_type"] == "custom" and not hp.get("attention_layer_size") and precision == "bf16"
                        and enc_depth % 64 == 0 and hp["decoder_units"] % 64 == 0 and enc_depth <= 2048 and hp["decoder_units"] <= 592)
        if (hp.get("attention_layer_size") or hp["attention_type"] in ("bahdanau_monotonic", "custom")) and not custom_fused:
            self._init_attention_layer(params, hp, enc_depth, precision, device, scope)  # fp32 step-kernel decoder only
            return
        self.precision = precision
        self.att = hp["attention_type"]
        dt = _lib.torch_dtype(precision)
        self.D, self.Ud, self.V, self.L = enc_depth, hp["decoder_units"], hp["target_vocab_size"], hp["decoder_layers"]
        D, Ud, V = self.D, self.Ud, self.V
        up = lambda a, t=dt: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device=device, dtype=t).contiguous()
        wm = np.asarray(params[f"{scope}/memory_layer/kernel"], np.float32)  # [D, Ud]
        n_pad = _round_up(Ud, 128) if precision == "bf16" else Ud
        wmt = np.zeros((n_pad, D), np.float32)
        wmt[:Ud] = wm.T
        self.w_mem_t = up(wmt)
        pre = f"{scope}/decoder/attention_wrapper"
        self.w_cell, self.b_cell, self.w_cell_tc = [], [], []
        # tensor-core decoder (decoder_tc.cu): bf16, D and Ud multiples of 64 (plas.h)
        self.tc = precision == "bf16" and D % 64 == 0 and Ud % 64 == 0 and D <= 2048 and Ud <= 2048
        self.w_query_tc = self.w_proj_pad = self.w_vw_t = None
        self.w_x_tc, self.w_h_tc = [None] * 4, [None] * 4
        for k in range(self.L):
            kern = np.asarray(params[f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell/kernel"], np.float32)
            bias = np.asarray(params[f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell/bias"], np.float32)
            if k == 0:
                kern = fold_embedding(params, hp, scope, kern)
                assert kern.shape == (V + D + Ud, 4 * Ud), kern.shape
                self.w_emb = up(packing.pack_unit_major(kern[:V], Ud))
                rows = kern[V:]
            else:
                assert kern.shape == (2 * Ud, 4 * Ud), kern.shape
                rows = kern
            packed = packing.pack_cell_bf16(rows, Ud) if precision == "bf16" else packing.pack_cell_f32(rows, Ud)
            self.w_cell.append(up(packed))
            if self.tc:
                self.w_cell_tc.append(up(packing.pack_cell_tc(rows, Ud)))
                # folded-context decoder (decoder_fold.cu): every cell contracts over h only (K = Ud); the context rows of
                # cell 0 move into VW = values . W0[V:V+D] (prepare_memory), 4Ud wide with unit-major columns
                if k == 0:
                    self.w_vw_t = up(np.ascontiguousarray(packing.pack_unit_major(rows[:D], Ud).T))  # [4Ud, D] K-major
                    self.w_h_tc[0] = up(packing.pack_cell_tc(rows[D:], Ud))
                else:
                    self.w_x_tc[k] = up(packing.pack_cell_tc(rows[:Ud], Ud))
                    self.w_h_tc[k] = up(packing.pack_cell_tc(rows[Ud:], Ud))
            self.b_cell.append(up(packing.pack_unit_major(bias, Ud), torch.float32))
        self.w_query = self.v_att = self.score_bias_dev = None
        self.score_bias = 0.0
        if self.att == "bahdanau":
            self.w_query = up(params[f"{pre}/bahdanau_attention/query_layer/kernel"])
            self.v_att = up(params[f"{pre}/bahdanau_attention/attention_v"], torch.float32)
            if self.tc:
                self.w_query_tc = up(packing.pack_query_tc(params[f"{pre}/bahdanau_attention/query_layer/kernel"], Ud))
        elif self.att == "custom":  # CustomAttention's own query layer (las/model.py:88-89), scope attention_wrapper/query_layer
            self.w_query = up(params[f"{pre}/query_layer/kernel"])
            self.w_query_tc = up(packing.pack_query_tc(params[f"{pre}/query_layer/kernel"], Ud))
        elif self.att == "luong_monotonic":
            self.score_bias = float(params[f"{pre}/luong_monotonic_attention/attention_score_bias"])
            self.score_bias_dev = torch.full((1,), self.score_bias, dtype=torch.float32, device=device)
        self.w_proj_t = up(np.asarray(params[f"{scope}/decoder/projection_layer/kernel"], np.float32).T)  # [V, D]
        self.b_proj = up(params[f"{scope}/decoder/projection_layer/bias"], torch.float32)
        # fp32: the step-kernel decoder (csrc/train_dec.cu, plas_decoder_infer_f32) reads the TF layout directly
        self.tf = None
        if precision == "fp32" and self.att in ("luong", "bahdanau", "luong_monotonic") and Ud % 16 == 0 and D % 4 == 0:
            self.tf = dict(
                kernel=[up(fold_embedding(params, hp, scope, np.asarray(params[f"{pre}/multi_rnn_cell/cell_0/lstm_cell/kernel"], np.float32)) if k == 0
                           else params[f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell/kernel"], torch.float32) for k in range(self.L)],
                bias=[up(params[f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell/bias"], torch.float32) for k in range(self.L)],
                w_proj=up(params[f"{scope}/decoder/projection_layer/kernel"], torch.float32))
        if self.tc:
            wp = np.zeros((_round_up(V, 128), D), np.float32)
            wp[:V] = np.asarray(params[f"{scope}/decoder/projection_layer/kernel"], np.float32).T
            self.w_proj_pad = up(wp)


def _init_bottom_only(self, params, hp, enc_depth, precision, device, scope):
    """GNMT-style AttentionMultiCell wiring (las/model.py:20-69, 185-193): fp32 step-kernel decoder only."""
    if precision != "fp32":
        raise NotImplementedError("--bottom_only is built for the fp32 step-kernel decoder (precision='fp32')")
    self.precision, self.att = precision, hp["attention_type"]
    if self.att not in _lib.ATT_CODES:
        raise NotImplementedError(f"--bottom_only with attention_type={self.att}")
    self.D, self.Ud, self.V, self.L = enc_depth, hp["decoder_units"], hp["target_vocab_size"], hp["decoder_layers"]
    D, Ud, V = self.D, self.Ud, self.V
    if Ud % 16 or D % 4:
        raise NotImplementedError("--bottom_only needs decoder_units % 16 == 0 and encoder depth % 4 == 0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device).contiguous()
    wm = np.asarray(params[f"{scope}/memory_layer/kernel"], np.float32)
    self.w_mem_t = up(wm.T)
    pre = f"{scope}/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper"
    names = [f"{pre}/lstm_cell"] + [f"{scope}/decoder/multi_rnn_cell/cell_{k}/lstm_cell" for k in range(1, self.L)]
    kernels = [np.asarray(params[n + "/kernel"], np.float32) for n in names]
    kernels[0] = fold_embedding(params, hp, scope, kernels[0])
    A = int(hp.get("attention_layer_size") or 0) or D  # the wrapped cell 0 emits Dense([h0; context]) when attention_layer_size is set
    if A % 4:
        raise NotImplementedError("attention_layer_size must be a multiple of 4")
    for k, kern in enumerate(kernels):
        din = (V + A) if k == 0 else ((A if k == 1 else Ud) + A)
        assert kern.shape == (din + Ud, 4 * Ud), (k, kern.shape)
    self.tf = dict(kernel=[up(k) for k in kernels], bias=[up(params[n + "/bias"]) for n in names],
                   w_proj=up(params[f"{scope}/decoder/projection_layer/kernel"]))
    if hp.get("attention_layer_size"):
        self.A = A
        self.tf["w_att_layer"] = up(params[f"{pre}/attention_layer/kernel"])
        assert self.tf["w_att_layer"].shape == (Ud + D, A)
    self.b_proj = up(params[f"{scope}/decoder/projection_layer/bias"])
    _attention_params(self, params, pre, up, device)
    self.tc = False


def _attention_params(self, params, pre, up, device):
    """Attention-mechanism variables of the fp32 step-kernel decoder.  Scopes as tf.contrib.seq2seq names them under the
    AttentionWrapper scope ``pre`` ([3P-recalled]): ``<type>_attention/{query_layer/kernel, attention_v,
    attention_score_bias}``; CustomAttention (las/model.py:72-101) calls its own Dense 'query_layer' outside the luong scope."""
    self.w_query = self.v_att = self.score_bias_dev = None
    self.score_bias = 0.0
    if self.att in ("bahdanau", "bahdanau_monotonic"):
        self.w_query = up(params[f"{pre}/{self.att}_attention/query_layer/kernel"])
        self.v_att = up(params[f"{pre}/{self.att}_attention/attention_v"])
    elif self.att == "custom":
        self.w_query = up(params[f"{pre}/query_layer/kernel"])
    if self.att.endswith("_monotonic"):
        self.score_bias = float(params[f"{pre}/{self.att}_attention/attention_score_bias"])
        self.score_bias_dev = torch.full((1,), self.score_bias, dtype=torch.float32, device=device)


def _init_attention_layer(self, params, hp, enc_depth, precision, device, scope, binf=None):
    """The default wiring on the fp32 step-kernel decoder only: attention_layer_size = A (las/model.py:180-200: AttentionWrapper's
    Dense over [cell output; context]; the attention fed back to cell 0 and read by the projection is A wide) and / or the
    attention types the fused decoders do not carry (bahdanau_monotonic, custom)."""
    if precision != "fp32":
        raise NotImplementedError("attention_layer_size / bahdanau_monotonic / custom attention are built for the fp32 step-kernel decoder")
    self.precision, self.att = precision, hp["attention_type"]
    self.bottom_only = self.pass_hidden_state = False
    self.D, self.Ud, self.V, self.L = enc_depth, hp["decoder_units"], hp["target_vocab_size"], hp["decoder_layers"]
    has_layer = bool(hp.get("attention_layer_size"))
    D, Ud, V = self.D, self.Ud, self.V
    A = int(hp["attention_layer_size"]) if has_layer else D
    if Ud % 16 or D % 4 or A % 4:
        raise NotImplementedError("the step-kernel decoder needs decoder_units % 16 == 0, encoder depth % 4 == 0 and A % 4 == 0")
    if has_layer:
        self.A = A
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device).contiguous()
    self.w_mem_t = up(np.asarray(params[f"{scope}/memory_layer/kernel"], np.float32).T)
    pre = f"{scope}/decoder/attention_wrapper"
    names = [f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell" for k in range(self.L)]
    kernels = [np.asarray(params[n + "/kernel"], np.float32) for n in names]
    if binf is None:
        kernels[0] = fold_embedding(params, hp, scope, kernels[0])
    if binf is not None:  # --binf_projection: fold embedding_fn and transform_binf_to_phones into the weights (see __init__)
        n = binf.shape[0]
        assert binf.shape[1] == V and A == 2 * n and kernels[0].shape == (n + A + Ud, 4 * Ud), (binf.shape, A, kernels[0].shape)
        kernels[0] = np.concatenate([binf.T.astype(np.float64) @ kernels[0][:n].astype(np.float64), kernels[0][n:]], 0).astype(np.float32)
        w_proj, b_proj = np.concatenate([binf, 1.0 - binf], 0), np.zeros((V,), np.float32)
    else:
        w_proj, b_proj = params[f"{scope}/decoder/projection_layer/kernel"], params[f"{scope}/decoder/projection_layer/bias"]
    assert kernels[0].shape == (V + A + Ud, 4 * Ud), kernels[0].shape
    self.tf = dict(kernel=[up(k) for k in kernels], bias=[up(params[n + "/bias"]) for n in names], w_proj=up(w_proj))
    if has_layer:
        self.tf["w_att_layer"] = up(params[f"{pre}/attention_layer/kernel"])
        assert self.tf["w_att_layer"].shape == (Ud + D, A)
    assert self.tf["w_proj"].shape == (A, V)
    self.b_proj = up(b_proj)
    _attention_params(self, params, pre, up, device)
    self.tc = False


def _train_state(self):
    """The variables as a ``train.TrainState`` (flat fp32 buffer in the TF layout) for the TRAIN-mode kernels; built on first use."""
    if self._train_state is None:
        from . import train as tr
        self._train_state = tr.TrainState({k: v for k, v in self._params.items() if not k.startswith("_")}, device=self._device)
    return self._train_state


SpellerWeights.train_state = _train_state
SpellerWeights._init_bottom_only = _init_bottom_only
SpellerWeights._init_attention_layer = _init_attention_layer


def prepare_memory(encoder_outputs, source_sequence_length, w, memory_is_masked=False):
    """values = length-masked memory; keys = memory_layer(values) (tf.contrib.seq2seq
    _BaseAttentionMechanism; reference las/model.py:168-169)."""
    L = _lib.lib()
    B, Tm, D = encoder_outputs.shape
    enc = encoder_outputs.contiguous()
    if memory_is_masked:
        values = enc  # already zero for t >= length (our listener guarantees it)
    else:
        values = torch.empty_like(enc)
        with _lib.stage("mask"):
            _lib.check(L.plas_mask_time(_lib.dtype_code(w.precision), _lib.ptr(enc), _lib.ptr(values),
                                        _lib.ptr(source_sequence_length), B, Tm, D, _lib.stream_ptr()))
        _lib.count_launches(1)
    n_pad = w.w_mem_t.shape[0]
    keys = torch.empty((B * Tm, n_pad), dtype=enc.dtype, device=enc.device)
    fn = L.plas_gemm_bf16 if w.precision == "bf16" else L.plas_gemm_f32
    with _lib.stage("memory_gemm"):
        _lib.check(fn(_lib.ptr(values), B * Tm, D, D, _lib.ptr(w.w_mem_t), n_pad, D, None, _lib.ptr(keys), n_pad,
                      _lib.stream_ptr()))
    _lib.count_launches(1)
    if n_pad != w.Ud:
        keys = keys[:, :w.Ud].contiguous()
    if w.att == "custom":  # CustomAttention: keys = relu(memory_layer(values)) (las/model.py:94)
        if keys.dtype == torch.float32:
            _lib.check(L.plas_relu_f32(_lib.ptr(keys), keys.numel(), _lib.stream_ptr()))
            _lib.count_launches(1)
        else:
            keys.relu_()  # bf16: exact, elementwise glue
    pv = vw = None
    if w.tc and B <= 128 and w.w_vw_t is not None and (4 * w.Ud) % 128 == 0:
        # VW = values x W0[V:V+D] (bf16): cell 0's share of the fed-back context, emitted by the attention phase as a . VW
        vw = torch.empty((B * Tm, 4 * w.Ud), dtype=enc.dtype, device=enc.device)
        with _lib.stage("memory_gemm"):
            _lib.check(L.plas_gemm_bf16(_lib.ptr(values), B * Tm, D, D, _lib.ptr(w.w_vw_t), 4 * w.Ud, D, None, _lib.ptr(vw),
                                        4 * w.Ud, _lib.stream_ptr()))
        _lib.count_launches(1)
    if w.tc and B <= 128:
        # PV = values x projection kernel (f32): the decoder forms logits as alignments . PV + bias
        vp = w.w_proj_pad.shape[0]
        pv = torch.empty((B * Tm, vp), dtype=torch.float32, device=enc.device)
        with _lib.stage("memory_gemm"):
            _lib.check(L.plas_gemm_bf16_f32out(_lib.ptr(values), B * Tm, D, D, _lib.ptr(w.w_proj_pad), vp, D, None,
                                               _lib.ptr(pv), vp, _lib.stream_ptr()))
        _lib.count_launches(1)
    return keys.view(B, Tm, w.Ud), values, pv, vw


def decode(encoder_outputs, source_sequence_length, w, hp, forced_ids=None, max_steps=None,
           want_alignment=True, trim=True, memory_is_masked=False, initial_state=None):
    """Run the decoder kernel.  Greedy when ``forced_ids`` is None (max_steps defaults to
    rint(Tm * decoding_length_factor), las/model.py:270-274); teacher-forced otherwise."""
    L = _lib.lib()
    dev = encoder_outputs.device
    B, Tm, D = encoder_outputs.shape
    assert D == w.D
    mem_len = source_sequence_length.to(device=dev, dtype=torch.int32).contiguous()
    keys, values, pv, vw = prepare_memory(encoder_outputs, mem_len, w, memory_is_masked)
    factor = float(hp.get("decoding_length_factor", 1.0))
    if forced_ids is not None:
        forced_ids = forced_ids.to(device=dev, dtype=torch.int32).contiguous()
        steps = forced_ids.shape[1] if max_steps is None else min(max_steps, forced_ids.shape[1])
        if forced_ids.shape[1] != steps:
            forced_ids = forced_ids[:, :steps].contiguous()
    else:
        steps = int(np.rint(np.float32(Tm) * np.float32(factor))) if max_steps is None else max_steps
    steps = max(int(steps), 0)
    cap = max(steps, 1)
    logits = torch.zeros((B, cap, w.V), dtype=torch.float32, device=dev)
    ids = torch.zeros((B, cap), dtype=torch.int32, device=dev)
    align = torch.zeros((B, cap, Tm), dtype=torch.float32, device=dev) if want_alignment else None
    seq_len = torch.zeros((B,), dtype=torch.int32, device=dev)
    n_steps = torch.zeros((1,), dtype=torch.int32, device=dev)
    if w.tf is not None and os.environ.get("PLAS_DEC_IMPL") != "simt":
        return _decode_f32_steps(L, w, hp, keys, values, mem_len, forced_ids, steps, cap, factor, logits, ids, align, seq_len,
                                 n_steps, trim, initial_state)
    d = _lib.DecDesc()
    d.dtype = _lib.dtype_code(w.precision)
    d.B, d.Tm, d.D, d.Ud, d.V, d.n_layers = B, Tm, D, w.Ud, w.V, w.L
    d.attention_type = _lib.ATT_CODES[w.att]
    d.sos_id, d.eos_id = hp["sos_id"], hp["eos_id"]
    d.max_steps = cap if steps > 0 else 0
    d.teacher_forced = 1 if forced_ids is not None else 0
    d.decoding_length_factor = factor
    d.score_bias = w.score_bias
    d.keys, d.values, d.mem_len = keys.data_ptr(), values.data_ptr(), mem_len.data_ptr()
    for k in range(w.L):
        d.w_cell[k] = w.w_cell[k].data_ptr()
        d.b_cell[k] = w.b_cell[k].data_ptr()
    d.w_emb = w.w_emb.data_ptr()
    d.w_query = w.w_query.data_ptr() if w.w_query is not None else None
    d.v_att = w.v_att.data_ptr() if w.v_att is not None else None
    d.w_proj, d.b_proj = w.w_proj_t.data_ptr(), w.b_proj.data_ptr()
    d.forced_ids = forced_ids.data_ptr() if forced_ids is not None else None
    d.logits, d.sample_ids = logits.data_ptr(), ids.data_ptr()
    d.alignment = align.data_ptr() if align is not None else None
    d.seq_len, d.n_steps = seq_len.data_ptr(), n_steps.data_ptr()
    if pv is not None:
        for k in range(w.L):
            d.w_cell_tc[k] = w.w_cell_tc[k].data_ptr()
        d.w_query_tc = w.w_query_tc.data_ptr() if w.w_query_tc is not None else None
        d.pv, d.pv_ld = pv.data_ptr(), pv.shape[1]
        if vw is not None:
            d.vw = vw.data_ptr()
            for k in range(w.L):
                d.w_h_tc[k] = w.w_h_tc[k].data_ptr()
                if k:
                    d.w_x_tc[k] = w.w_x_tc[k].data_ptr()
    need = L.plas_decoder_workspace_bytes(C.byref(d))
    ws = torch.empty((need,), dtype=torch.uint8, device=dev)
    with _lib.grid_sync_kernel(), _lib.stage("decoder"):
        _lib.check(L.plas_decoder_fwd(C.byref(d), _lib.ptr(ws), need, _lib.stream_ptr()))
    _lib.count_launches(1)
    if trim:
        n = int(n_steps.item()) if steps > 0 else 0  # device->host read of the step count
        logits, ids = logits[:, :n], ids[:, :n]
        if align is not None:
            align = align[:, :n]
    return logits, ids, align, seq_len, n_steps


def _decode_f32_steps(L, w, hp, keys, values, mem_len, forced_ids, steps, cap, factor, logits, ids, align, seq_len, n_steps, trim,
                      initial_state=None, beam=None):
    """fp32 (reference-precision) decode through plas_decoder_infer_f32: a loop of step kernels on the TF weight layout."""
    B, Tm, D = values.shape
    d = _lib.DecInferDesc()
    d.B, d.Tm, d.D, d.Ud, d.V, d.n_layers = B, Tm, D, w.Ud, w.V, w.L
    d.attention_type = _lib.ATT_CODES[w.att]
    d.sos_id, d.eos_id = hp["sos_id"], hp["eos_id"]
    d.max_steps = cap if steps > 0 else 0
    d.teacher_forced = 1 if forced_ids is not None else 0
    d.decoding_length_factor = factor
    for k in range(w.L):
        d.kernel[k], d.bias[k] = w.tf["kernel"][k].data_ptr(), w.tf["bias"][k].data_ptr()
    d.w_query = w.w_query.data_ptr() if w.w_query is not None else None
    d.v_att = w.v_att.data_ptr() if w.v_att is not None else None
    d.w_proj, d.b_proj = w.tf["w_proj"].data_ptr(), w.b_proj.data_ptr()
    d.score_bias = w.score_bias_dev.data_ptr() if w.score_bias_dev is not None else None
    d.keys, d.values, d.mem_len = keys.data_ptr(), values.data_ptr(), mem_len.data_ptr()
    d.forced_ids = forced_ids.data_ptr() if forced_ids is not None else None
    d.logits = logits.data_ptr() if logits is not None else None
    d.sample_ids = ids.data_ptr() if ids is not None else None
    d.alignment = align.data_ptr() if align is not None else None
    if beam is not None:  # BeamSearchDecoder on the tiled batch
        d.beam_width = beam["width"]
        d.beam_predicted, d.beam_parent, d.beam_word = beam["predicted"].data_ptr(), beam["parent"].data_ptr(), beam["word"].data_ptr()
        d.beam_scores, d.beam_lengths = beam["scores"].data_ptr(), beam["lengths"].data_ptr()
    d.seq_len, d.n_steps = seq_len.data_ptr(), n_steps.data_ptr()
    d.bottom_only = 1 if getattr(w, "bottom_only", False) else 0
    if "w_att_layer" in w.tf:
        d.att_layer, d.w_att_layer = w.A, w.tf["w_att_layer"].data_ptr()
    keep_alive = []
    if initial_state is not None:  # pass_hidden_state: cell l starts from (c, h) number l of the listener's final state
        for l, (c0, h0) in enumerate(initial_state[:w.L]):
            c0, h0 = c0.to(torch.float32).contiguous(), h0.to(torch.float32).contiguous()
            assert c0.shape == (B, w.Ud), "pass_hidden_state needs encoder_units == decoder_units"
            keep_alive += [c0, h0]
            d.c_init[l], d.h_init[l] = c0.data_ptr(), h0.data_ptr()
    need = L.plas_decoder_infer_f32_workspace_bytes(C.byref(d))
    ws = torch.empty((need,), dtype=torch.uint8, device=values.device)
    with _lib.stage("decoder"):
        _lib.check(L.plas_decoder_infer_f32(C.byref(d), _lib.ptr(ws), need, _lib.stream_ptr()))
    _lib.count_launches(1 + d.max_steps * (4 + w.L))
    if beam is not None:
        return None
    if trim:
        n = int(n_steps.item()) if steps > 0 else 0
        logits, ids = logits[:, :n], ids[:, :n]
        if align is not None:
            align = align[:, :n]
    return logits, ids, align, seq_len, n_steps


FinalBeamSearchDecoderOutput = namedtuple("FinalBeamSearchDecoderOutput", ["predicted_ids", "beam_search_decoder_output"])
BeamSearchDecoderOutput = namedtuple("BeamSearchDecoderOutput", ["scores", "predicted_ids", "parent_ids"])


def decode_beam(encoder_outputs, source_sequence_length, w, hp, beam_width, memory_is_masked=False, initial_state=None):
    """PREDICT with beam_width > 0 (las/model.py:215-226,298-319): tile_batch of the memory (and of the initial state),
    tf.contrib.seq2seq.BeamSearchDecoder from the sos token with length penalty 0, dynamic_decode, gather_tree.  fp32 step-kernel
    decoder.  Returns (FinalBeamSearchDecoderOutput(predicted_ids [B, T, W], ...), final log-probs [B, W], lengths [B, W],
    sequence lengths [B, W], n_steps)."""
    if w.tf is None:
        raise NotImplementedError("beam search is built on the fp32 step-kernel decoder (precision='fp32')")
    L = _lib.lib()
    W = int(beam_width)
    dev = encoder_outputs.device
    Bb, Tm, D = encoder_outputs.shape
    enc = encoder_outputs.repeat_interleave(W, dim=0).contiguous()  # tile_batch: row b*W + w
    mem_len = source_sequence_length.to(device=dev, dtype=torch.int32).repeat_interleave(W).contiguous()
    if initial_state is not None:
        initial_state = [(c.repeat_interleave(W, dim=0), h.repeat_interleave(W, dim=0)) for c, h in initial_state]
    keys, values, _, _ = prepare_memory(enc, mem_len, w, memory_is_masked)
    factor = float(hp.get("decoding_length_factor", 1.0))
    steps = max(int(np.rint(np.float32(Tm) * np.float32(factor))), 0)
    cap = max(steps, 1)
    B = Bb * W
    beam = dict(width=W, predicted=torch.full((Bb, cap, W), int(hp["eos_id"]), dtype=torch.int32, device=dev),
                parent=torch.zeros((cap, B), dtype=torch.int32, device=dev), word=torch.zeros((cap, B), dtype=torch.int32, device=dev),
                scores=torch.zeros((B,), dtype=torch.float32, device=dev), lengths=torch.zeros((B,), dtype=torch.int32, device=dev))
    seq_len = torch.zeros((B,), dtype=torch.int32, device=dev)
    n_steps = torch.zeros((1,), dtype=torch.int32, device=dev)
    _decode_f32_steps(L, w, hp, keys, values, mem_len, None, steps, cap, factor, None, None, None, seq_len, n_steps, False,
                      initial_state, beam=beam)
    n = int(n_steps.item()) if steps > 0 else 0
    step_out = BeamSearchDecoderOutput(None, beam["word"][:n].view(n, Bb, W).permute(1, 0, 2), beam["parent"][:n].view(n, Bb, W).permute(1, 0, 2))
    out = FinalBeamSearchDecoderOutput(beam["predicted"][:, :n], step_out)
    return out, beam["scores"].view(Bb, W), beam["lengths"].view(Bb, W), seq_len.view(Bb, W), n_steps


def _speller_train_stochastic(encoder_outputs, decoder_inputs, source_sequence_length, target_sequence_length, hparams, weights, init,
                              memory_is_masked, step):
    """The TRAIN branch with its RNG-driven parts (las/model.py:276-296): scheduled sampling (sampling_probability > 0: after step
    t a Bernoulli(p) draw decides per utterance whether step t+1 is fed a sample from Categorical(logits_t) instead of the teacher
    id) and DropoutWrapper input dropout (las/ops.py:14-18).  Runs the forward pass of the training kernels (train.SpellerTrain;
    fp32; counter-based RNG keyed by hparams['dropout_seed'] and the optimiser ``step``, mirrored in numpy by
    train.reference_sampling / reference_masks).  sample_id follows the scheduled helpers: the id drawn at step t where the row
    sampled, -1 otherwise (a draw that equals the teacher id cannot be told apart and reads -1)."""
    from . import train as tr
    if weights.precision != "fp32" or getattr(weights, "binf_projection", False):
        raise NotImplementedError("TRAIN-mode dropout / scheduled sampling run on the fp32 training kernels of the phone speller")
    L = _lib.lib()
    st = weights.train_state()
    if step is not None:
        st.step = int(step)
    dev = encoder_outputs.device
    B, Tm, D = encoder_outputs.shape
    V = hparams["target_vocab_size"]
    steps = decoder_inputs.shape[1]
    if target_sequence_length is not None:
        steps = min(steps, int(target_sequence_length.max().item()))
        if hparams.get("max_symbols", -1) and hparams.get("max_symbols", -1) > 0:
            steps = min(steps, hparams["max_symbols"])
    ids = decoder_inputs[:, :steps].to(device=dev, dtype=torch.int64).contiguous()
    mem_len = source_sequence_length.to(device=dev, dtype=torch.int32).contiguous()
    memory = encoder_outputs.to(torch.float32).contiguous()
    if not memory_is_masked:
        masked = torch.empty_like(memory)
        _lib.check(L.plas_mask_time(_lib.dtype_code("fp32"), _lib.ptr(memory), _lib.ptr(masked), _lib.ptr(mem_len), B, Tm, D, _lib.stream_ptr()))
        _lib.count_launches(1)
        memory = masked
    emb = bool(hparams.get("embedding_size"))
    table = st.view(f"{weights.scope}/target_embedding") if emb else None
    x_in = table[ids].contiguous() if emb else torch.nn.functional.one_hot(ids, V).to(torch.float32)
    sp = tr.SpellerTrain(st, hparams, weights.scope, x_in.shape[2], V, index=0, table=table)
    logits = sp.forward(memory, mem_len, x_in, initial_state=init, ids=ids)
    sample_id = torch.full((B, steps), -1, dtype=torch.int32, device=dev)
    if sp.fed_ids is not None and steps > 1:
        fed, teach = sp.fed_ids[:, 1:steps].to(torch.int32), ids[:, 1:].to(torch.int32)
        sample_id[:, :steps - 1] = torch.where(fed != teach, fed, torch.full_like(fed, -1))
    elif float(hparams.get("sampling_probability", 0.0)) == 0.0:
        sample_id = logits.argmax(-1).to(torch.int32)  # TrainingHelper
    n_steps = torch.full((1,), steps, dtype=torch.int32, device=dev)
    return BasicDecoderOutput(logits, sample_id), SpellerState(None, n_steps), target_sequence_length


def speller(encoder_outputs, encoder_state, decoder_inputs, source_sequence_length, target_sequence_length,
            mode, hparams, weights, binary_outputs=False, binf_embedding=None, transparent_projection=False,
            memory_is_masked=False, want_alignment=True, trim=True, step=None):
    """las/model.py:205-349.  mode 'train'/'eval' with ``decoder_inputs`` (targets_inputs ids [B,L]) runs teacher forcing
    (TrainingHelper), in 'train' with scheduled sampling when hparams['sampling_probability'] > 0 and input dropout when
    hparams['dropout'] > 0 (``step`` selects the RNG stream, as in train.train_step); otherwise greedy."""
    if binf_embedding is not None and not getattr(weights, "binf_projection", False):
        raise NotImplementedError("a binf2phone matrix is used by the --binf_projection wiring only: build the SpellerWeights with binf=")
    if binary_outputs or transparent_projection:
        # binary_outputs without --binf_projection: the reference's own non-TRAIN graph slices an n-wide output as [n:2n]
        # (utils/training_helper.py:19-21) and cannot be built; with --binf_sampling it decodes through InferenceHelper
        raise NotImplementedError("binary-feature decoding is built for --binf_projection only (DESIGN.md)")
    init = None
    if getattr(weights, "pass_hidden_state", False):  # las/model.py:259-267
        init = encoder_state if isinstance(encoder_state[0], (tuple, list)) else (encoder_state,)
    if mode == "train" and (float(hparams.get("sampling_probability", 0.0)) > 0.0 or float(hparams.get("dropout", 0.0)) > 0.0):
        return _speller_train_stochastic(encoder_outputs, decoder_inputs, source_sequence_length, target_sequence_length, hparams,
                                         weights, init, memory_is_masked, step)
    if mode == "train":
        steps = None
        if target_sequence_length is not None:
            steps = int(target_sequence_length.max().item())
            if hparams.get("max_symbols", -1) and hparams.get("max_symbols", -1) > 0:
                steps = min(steps, hparams["max_symbols"])
        logits, ids, align, seq_len, n_steps = decode(encoder_outputs, source_sequence_length, weights, hparams,
                                                      forced_ids=decoder_inputs, max_steps=steps,
                                                      want_alignment=False, memory_is_masked=memory_is_masked,
                                                      initial_state=init)
        seq_len = target_sequence_length
    elif mode == "infer" and int(hparams.get("beam_width", 0) or 0) > 0:  # las/model.py:215-226,298-319 (PREDICT only)
        out, scores, lengths, seq_len, n_steps = decode_beam(encoder_outputs, source_sequence_length, weights, hparams,
                                                             int(hparams["beam_width"]), memory_is_masked, init)
        return out, SpellerState(None, n_steps), seq_len
    else:
        logits, ids, align, seq_len, n_steps = decode(encoder_outputs, source_sequence_length, weights, hparams,
                                                      memory_is_masked=memory_is_masked, want_alignment=want_alignment,
                                                      trim=trim, initial_state=init)
    return BasicDecoderOutput(logits, ids), SpellerState(align, n_steps), seq_len
