"""Listener operator: pyramidal (bi)directional LSTM encoder on the GPU.

Mirrors ``las.model.listener`` / ``las.ops.pyramidal_bilstm`` (reference las/model.py:104-142,
las/ops.py:23-87): same inputs (``encoder_inputs [B,T,C]``, ``source_sequence_length [B]``, mode,
encoder hparams), same outputs ``((outputs [B,T',D], lengths [B]), state)``.  Per layer the work is
  K2  xproj = x @ W_x + b          one time-parallel GEMM for both directions (plas_gemm_*)
  K3  persistent recurrence         (plas_bilstm_rec_fwd)
and the pyramidal frame concat (las/ops.py:49-65) is a free view: layer outputs are allocated with
an even, zero-padded time extent so ``[B,T,2U] -> [B,T/2,4U]`` is a reshape of the same bytes.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, packing


def _round_up(x, m):
    return (x + m - 1) // m * m


class ListenerWeights:
    """Device-resident, kernel-layout copy of the ``listener/`` variables."""

    def __init__(self, params, hp, num_channels, precision="fp32", device="cuda"):
        _lib.require_cuda()
        self._params, self._device, self._train_state = params, device, None  # TRAIN mode with dropout: train_state()
        self.pyramidal = bool(hp["use_pyramidal"])
        self.precision = precision
        self.U = U = hp["encoder_units"]
        self.L = hp["encoder_layers"]
        self.ndir = 1 if hp["unidirectional"] else 2
        self.C = num_channels
        dt = _lib.torch_dtype(precision)
        self.upc = _lib.lib().plas_rec_units_per_cta(_lib.dtype_code(precision), U)
        self.layers = []
        din = num_channels
        for l in range(self.L):
            if not self.pyramidal:  # stacked MultiRNNCell per direction (las/model.py:111-142)
                dirs = ["bidirectional_rnn/fw", "bidirectional_rnn/bw"] if self.ndir == 2 else ["rnn"]
                names = [f"listener/{d}/multi_rnn_cell/cell_{l}/lstm_cell" for d in dirs]
            elif self.ndir == 2:
                names = [f"listener/bilstm_{l}/bidirectional_rnn/{d}/lstm_cell" for d in ("fw", "bw")]
            else:
                names = [f"listener/bilstm_{l}/rnn/lstm_cell"]
            kernels = [np.asarray(params[n + "/kernel"], np.float32) for n in names]
            biases = [np.asarray(params[n + "/bias"], np.float32) for n in names]
            assert kernels[0].shape == (din + U, 4 * U), (kernels[0].shape, din, U)
            k_pad = _round_up(din, 64) if precision == "bf16" else din
            wt, bs = packing.pack_inproj(kernels, biases, din, U, k_pad)
            # non-pyramidal layers >= 1: direction d reads only its own previous output, so the projection is one
            # GEMM per direction (rows d*4U.. of wt are that direction's weights over its U inputs)
            if precision == "bf16":
                whh = packing.pack_rec_bf16(kernels, din, U)
            else:
                whh = packing.pack_rec_f32(kernels, din, U, self.upc)
            whh_tc = None
            if precision == "bf16" and U in (64, 128, 256, 512):
                whh_tc = torch.from_numpy(packing.pack_rec_tc(kernels, din, U)).to(device=device, dtype=dt).contiguous()
            # fp32: the row-blocked recurrence of csrc/train_rec.cu (forward-only call) reads the TF layout directly
            tf_kernel = [torch.from_numpy(k).to(device) for k in kernels] if precision == "fp32" and U % 4 == 0 else None
            tf_bias = [torch.from_numpy(b).to(device) for b in biases] if tf_kernel is not None else None
            self.layers.append(dict(
                din=din, k_pad=k_pad, whh_tc=whh_tc, tf_kernel=tf_kernel, tf_bias=tf_bias,
                wt=torch.from_numpy(wt).to(device=device, dtype=dt).contiguous(),
                bias=torch.from_numpy(bs).to(device),
                whh=torch.from_numpy(whh).to(device=device, dtype=dt).contiguous()))
            din = (self.ndir * U * (1 if l == 0 else 2)) if self.pyramidal else U
        self.out_depth = (self.ndir * U * (2 if self.L > 1 else 1)) if self.pyramidal else self.ndir * U


def _train_state(self):
    """The variables as a ``train.TrainState`` (flat fp32 buffer in the TF layout) for the TRAIN-mode kernels; built on first use."""
    if self._train_state is None:
        from . import train as tr
        self._train_state = tr.TrainState({k: v for k, v in self._params.items() if not k.startswith("_")}, device=self._device)
    return self._train_state


ListenerWeights.train_state = _train_state


def _gemm(precision, a2d, M, K, lda, wt, bias, out2d):
    L = _lib.lib()
    N = wt.shape[0]
    fn = L.plas_gemm_bf16 if precision == "bf16" else L.plas_gemm_f32
    with _lib.stage("inproj_gemm"):
        _lib.check(fn(_lib.ptr(a2d), M, K, lda, _lib.ptr(wt), N, wt.stride(0),
                      _lib.ptr(bias) if bias is not None else None, _lib.ptr(out2d), out2d.stride(0),
                      _lib.stream_ptr()))
    _lib.count_launches(1)


def _gemm_view(precision, a, a_col0, M, K, lda, wt, w_row0, N, bias, out2d, out_col0):
    """C[:, out_col0:out_col0+N] = A[:, a_col0:a_col0+K] @ wt[w_row0:w_row0+N, :K]^T + bias[w_row0:...] on strided
    views of contiguous tensors (pointer arithmetic; every offset keeps the 16-byte alignment the kernels need)."""
    L = _lib.lib()
    esz = a.element_size()
    fn = L.plas_gemm_bf16 if precision == "bf16" else L.plas_gemm_f32
    with _lib.stage("inproj_gemm"):
        _lib.check(fn(C.c_void_p(a.data_ptr() + a_col0 * esz), M, K, lda,
                      C.c_void_p(wt.data_ptr() + w_row0 * wt.stride(0) * wt.element_size()), N, wt.stride(0),
                      C.c_void_p(bias.data_ptr() + w_row0 * 4), C.c_void_p(out2d.data_ptr() + out_col0 * out2d.element_size()),
                      out2d.stride(0), _lib.stream_ptr()))
    _lib.count_launches(1)


def rec_cluster_budget(B, U, ndir, sm_budget):
    """plas_rec_desc.max_clusters for a batch of B utterances under the pipelined loops' SM budget (``_lib.rec_sms``; 0 = none):
    the budget in clusters of U/32 CTAs -- unless it would push the plan beyond two 16-row groups per cluster: four 16-row groups
    on 4 clusters (2.94 us per step at B = 128) lose more than the other batch's GEMMs gain (three 15-row groups on 6: 2.11 us)."""
    if not sm_budget:
        return 0
    mc = max(ndir, sm_budget // max(1, U // 32))
    return mc if -(-B // (max(1, mc // ndir) * 2)) <= 16 else 0


def bilstm_layer(x, lengths, lw, U, ndir, precision, t_alloc_out, per_direction_input=False):
    """One (bi)LSTM layer: x [B,T,K] (compute dtype, K == lw['k_pad']) -> out [B,t_alloc_out,ndir*U].
    ``per_direction_input``: x is [B,T,ndir*U] and direction d reads only x[..., d*U:(d+1)*U] (stacked MultiRNNCell)."""
    L = _lib.lib()
    B, T, K = x.shape
    dt = x.dtype
    if lw.get("tf_kernel") is not None and os.environ.get("PLAS_REC_IMPL") != "l2":
        return _bilstm_layer_f32(x, lengths, lw, U, ndir, t_alloc_out, per_direction_input)
    xproj = torch.empty((B * T, ndir * 4 * U), dtype=dt, device=x.device)
    if per_direction_input:
        for dd in range(ndir):
            _gemm_view(precision, x, dd * U, B * T, U, K, lw["wt"], dd * 4 * U, 4 * U, lw["bias"], xproj, dd * 4 * U)
    else:
        _gemm(precision, x.reshape(B * T, K), B * T, K, K, lw["wt"], lw["bias"], xproj)
    out = torch.empty((B, t_alloc_out, ndir * U), dtype=dt, device=x.device)  # the kernel zero-fills t >= len
    c_fin = torch.empty((ndir, B, U), dtype=torch.float32, device=x.device)
    h_fin = torch.empty((ndir, B, U), dtype=torch.float32, device=x.device)
    d = _lib.RecDesc()
    d.dtype, d.B, d.T, d.U, d.ndir = _lib.dtype_code(precision), B, T, U, ndir
    d.out_zeroed = 0
    d.xproj, d.whh, d.lengths, d.out = xproj.data_ptr(), lw["whh"].data_ptr(), lengths.data_ptr(), out.data_ptr()
    d.out_batch_stride = out.stride(0)
    d.c_final, d.h_final = c_fin.data_ptr(), h_fin.data_ptr()
    d.whh_tc = lw["whh_tc"].data_ptr() if lw.get("whh_tc") is not None else None
    d.max_clusters = rec_cluster_budget(B, U, ndir, _lib.rec_sm_budget)
    need = L.plas_rec_workspace_bytes(C.byref(d))
    ws = torch.empty((need,), dtype=torch.uint8, device=x.device)
    # the tensor-core recurrence only synchronises inside its clusters; the cooperative fallbacks exchange h through L2 across
    # the whole grid and must not share the GPU with another grid-synchronising kernel (_lib.grid_sync_kernel)
    cluster_only = precision == "bf16" and lw.get("whh_tc") is not None and os.environ.get("PLAS_REC_IMPL", "tc") == "tc"
    if cluster_only:
        with _lib.stage("rec"):
            _lib.check(L.plas_bilstm_rec_fwd(C.byref(d), _lib.ptr(ws), need, _lib.stream_ptr()))
    else:
        with _lib.grid_sync_kernel(), _lib.stage("rec"):
            _lib.check(L.plas_bilstm_rec_fwd(C.byref(d), _lib.ptr(ws), need, _lib.stream_ptr()))
    _lib.count_launches(1)
    return out, (c_fin, h_fin)


def _bilstm_layer_f32(x, lengths, lw, U, ndir, t_alloc_out, per_direction_input):
    """fp32 (reference-precision) layer on the TF weight layout: strided fp32 GEMM per direction (plas_gemm_f32_ex) + the
    row-blocked persistent recurrence (plas_bilstm_rec_train_fwd called forward-only: nothing saved, final states returned)."""
    from .train import gemm_ex
    L = _lib.lib()
    B, T, K = x.shape
    din = U if per_direction_input else K
    z = torch.empty((B, T, ndir, 4 * U), dtype=torch.float32, device=x.device)
    # exact-fp32 SIMT GEMM on purpose: the 3xTF32 tensor-core GEMM of the training path (train.gemm_tc) is 2^-21-accurate per
    # product, which after three recurrent layers measures 1.4e-5 on encoder_out -- above this mode's 1e-5 bar
    with _lib.stage("inproj_gemm"):
        for dd in range(ndir):
            gemm_ex(B * T, 4 * U, din, x.data_ptr() + (4 * dd * U if per_direction_input else 0), K, 1, lw["tf_kernel"][dd].data_ptr(),
                    4 * U, 1, z.data_ptr() + 4 * dd * 4 * U, ndir * 4 * U, bias=lw["tf_bias"][dd].data_ptr())
    out = torch.zeros((B, t_alloc_out, ndir * U), dtype=torch.float32, device=x.device)
    c_fin = torch.empty((ndir, B, U), dtype=torch.float32, device=x.device)
    h_fin = torch.empty((ndir, B, U), dtype=torch.float32, device=x.device)
    d = _lib.RecTrainDesc()
    d.B, d.T, d.U, d.ndir, d.din = B, T, U, ndir, din
    d.z = z.data_ptr()
    for dd in range(ndir):
        d.kernel[dd] = lw["tf_kernel"][dd].data_ptr()
    d.lengths, d.out, d.out_batch_stride = lengths.data_ptr(), out.data_ptr(), out.stride(0)
    d.c_final, d.h_final = c_fin.data_ptr(), h_fin.data_ptr()
    need = L.plas_rec_train_workspace_bytes(C.byref(d))
    ws = torch.empty((need,), dtype=torch.uint8, device=x.device)
    with _lib.grid_sync_kernel(), _lib.stage("rec"):  # groups of CTAs exchange h through L2 when they are not one cluster
        _lib.check(L.plas_bilstm_rec_train_fwd(C.byref(d), _lib.ptr(ws), need, _lib.stream_ptr()))
    _lib.count_launches(1)
    return out, (c_fin, h_fin)


def pyramidal_bilstm(inputs, sequence_length, mode, hparams, weights):
    """las/ops.py:68-87 on the device.  ``inputs`` float32 [B,T,C]; returns outputs in the compute
    dtype of ``weights`` ([B,T',D]), the reduced lengths and the last layer's final state."""
    w = weights
    precision = w.precision
    B, T, Cin = inputs.shape
    assert Cin == w.C, f"features have {Cin} channels, weights expect {w.C}"
    lengths = sequence_length.to(device=inputs.device, dtype=torch.int32).contiguous()
    L = _lib.lib()
    if precision == "bf16":
        k_pad = w.layers[0]["k_pad"]
        x = torch.empty((B, T, k_pad), dtype=torch.bfloat16, device=inputs.device)
        src = inputs.contiguous()
        with _lib.stage("cast"):
            _lib.check(L.plas_cast_pad_bf16(_lib.ptr(src), B * T, Cin, Cin, _lib.ptr(x), k_pad, _lib.stream_ptr()))
        _lib.count_launches(1)
    else:
        x = inputs.to(torch.float32).contiguous()
    state = None
    for l, lw in enumerate(w.layers):
        T_l = x.shape[1]
        t_alloc = T_l if l == 0 else T_l + (T_l % 2)
        out, state = bilstm_layer(x, lengths, lw, w.U, w.ndir, precision, t_alloc)
        if l != 0:  # pyramidal_stack: free view + ceil-halved lengths (las/ops.py:49-65)
            out = out.view(B, t_alloc // 2, 2 * w.ndir * w.U)
            lengths = torch.div(lengths, 2, rounding_mode="floor") + lengths % 2
        x = out
    c_fin, h_fin = state
    if w.ndir == 2:
        enc_state = ((c_fin[0], h_fin[0]), (c_fin[1], h_fin[1]))
    else:
        enc_state = (c_fin[0], h_fin[0])
    return (x, lengths), enc_state


def stacked_bilstm(inputs, sequence_length, mode, hparams, weights):
    """Non-pyramidal branch of las/model.py:111-142: one MultiRNNCell stack per direction, no time reduction.
    Layer 0 projects the shared input for both directions in one GEMM; deeper layers project each direction's own
    previous output.  Returns outputs [B,T,ndir*U], the unchanged lengths and per-direction tuples of (c,h)."""
    w = weights
    precision = w.precision
    B, T, Cin = inputs.shape
    assert Cin == w.C, f"features have {Cin} channels, weights expect {w.C}"
    lengths = sequence_length.to(device=inputs.device, dtype=torch.int32).contiguous()
    L = _lib.lib()
    if precision == "bf16":
        k_pad = w.layers[0]["k_pad"]
        x = torch.empty((B, T, k_pad), dtype=torch.bfloat16, device=inputs.device)
        src = inputs.contiguous()
        with _lib.stage("cast"):
            _lib.check(L.plas_cast_pad_bf16(_lib.ptr(src), B * T, Cin, Cin, _lib.ptr(x), k_pad, _lib.stream_ptr()))
        _lib.count_launches(1)
    else:
        x = inputs.to(torch.float32).contiguous()
    states = []
    for l, lw in enumerate(w.layers):
        x, st = bilstm_layer(x, lengths, lw, w.U, w.ndir, precision, T, per_direction_input=(l > 0))
        states.append(st)
    per_dir = tuple(tuple((c[dd], h[dd]) for (c, h) in states) for dd in range(w.ndir))
    return (x, lengths), (per_dir if w.ndir == 2 else per_dir[0])


def _listener_train_dropout(encoder_inputs, source_sequence_length, hparams, weights, step):
    """TRAIN mode with input dropout (las/ops.py:14-18: DropoutWrapper(input_keep_prob = 1 - dropout) around every cell): the
    forward pass of the training path's kernels (train.listener_train_fwd; fp32, counter-based masks keyed by
    hparams['dropout_seed'] and the optimiser ``step``), returned in the operator's output format."""
    from . import train as tr
    if weights.precision != "fp32":
        raise NotImplementedError("TRAIN-mode dropout runs on the fp32 training kernels: build the ListenerWeights with precision='fp32'")
    st = weights.train_state()
    if step is not None:
        st.step = int(step)
    out, lengths, tape = tr.listener_train_fwd(encoder_inputs, source_sequence_length, st, hparams)
    ndir = tape[-1]["c_fin"].shape[0]
    if weights.pyramidal:  # final state of the last layer: (fw, bw) or the single direction's (c, h)
        per_dir = tuple((tape[-1]["c_fin"][dd], tape[-1]["h_fin"][dd]) for dd in range(ndir))
        return (out, lengths), (per_dir if ndir == 2 else per_dir[0])
    per_dir = tuple(tuple((t["c_fin"][dd], t["h_fin"][dd]) for t in tape) for dd in range(ndir))
    return (out, lengths), (per_dir if ndir == 2 else per_dir[0])


def listener(encoder_inputs, source_sequence_length, mode, hparams, weights, step=None):
    """las/model.py:104-142.  ``hparams`` is the flat dict (or the encoder view); ``weights`` a
    :class:`ListenerWeights`.  mode: 'train' | 'eval' | 'infer'.  In 'train' with hparams['dropout'] > 0 every LSTM cell
    input is dropped out (las/ops.py:14-18); ``step`` selects the mask stream (the optimiser step of train.train_step)."""
    if mode == "train" and float(hparams.get("dropout", 0.0)) > 0.0:
        return _listener_train_dropout(encoder_inputs, source_sequence_length, hparams, weights, step)
    if not weights.pyramidal:
        return stacked_bilstm(encoder_inputs, source_sequence_length, mode, hparams, weights)
    return pyramidal_bilstm(encoder_inputs, source_sequence_length, mode, hparams, weights)
