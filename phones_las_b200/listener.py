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

import numpy as np
import torch

from . import _lib, packing


def _round_up(x, m):
    return (x + m - 1) // m * m


class ListenerWeights:
    """Device-resident, kernel-layout copy of the ``listener/`` variables."""

    def __init__(self, params, hp, num_channels, precision="fp32", device="cuda"):
        _lib.require_cuda()
        if not hp["use_pyramidal"]:
            raise NotImplementedError("non-pyramidal listener (las/model.py:111-142) is not built yet")
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
            if self.ndir == 2:
                names = [f"listener/bilstm_{l}/bidirectional_rnn/{d}/lstm_cell" for d in ("fw", "bw")]
            else:
                names = [f"listener/bilstm_{l}/rnn/lstm_cell"]
            kernels = [np.asarray(params[n + "/kernel"], np.float32) for n in names]
            biases = [np.asarray(params[n + "/bias"], np.float32) for n in names]
            assert kernels[0].shape == (din + U, 4 * U), (kernels[0].shape, din, U)
            k_pad = _round_up(din, 64) if precision == "bf16" else din
            wt, bs = packing.pack_inproj(kernels, biases, din, U, k_pad)
            if precision == "bf16":
                whh = packing.pack_rec_bf16(kernels, din, U)
            else:
                whh = packing.pack_rec_f32(kernels, din, U, self.upc)
            whh_tc = None
            if precision == "bf16" and U in (64, 128, 256, 512):
                whh_tc = torch.from_numpy(packing.pack_rec_tc(kernels, din, U)).to(device=device, dtype=dt).contiguous()
            self.layers.append(dict(
                din=din, k_pad=k_pad, whh_tc=whh_tc,
                wt=torch.from_numpy(wt).to(device=device, dtype=dt).contiguous(),
                bias=torch.from_numpy(bs).to(device),
                whh=torch.from_numpy(whh).to(device=device, dtype=dt).contiguous()))
            din = self.ndir * U * (1 if l == 0 else 2)
        self.out_depth = din if self.L > 1 else self.ndir * U


def _gemm(precision, a2d, M, K, lda, wt, bias, out2d):
    L = _lib.lib()
    N = wt.shape[0]
    fn = L.plas_gemm_bf16 if precision == "bf16" else L.plas_gemm_f32
    with _lib.stage("inproj_gemm"):
        _lib.check(fn(_lib.ptr(a2d), M, K, lda, _lib.ptr(wt), N, wt.stride(0),
                      _lib.ptr(bias) if bias is not None else None, _lib.ptr(out2d), out2d.stride(0),
                      _lib.stream_ptr()))
    _lib.count_launches(1)


def bilstm_layer(x, lengths, lw, U, ndir, precision, t_alloc_out):
    """One (bi)LSTM layer: x [B,T,K] (compute dtype, K == lw['k_pad']) -> out [B,t_alloc_out,ndir*U]."""
    L = _lib.lib()
    B, T, K = x.shape
    dt = x.dtype
    xproj = torch.empty((B * T, ndir * 4 * U), dtype=dt, device=x.device)
    _gemm(precision, x.reshape(B * T, K), B * T, K, K, lw["wt"], lw["bias"], xproj)
    out = torch.zeros((B, t_alloc_out, ndir * U), dtype=dt, device=x.device)
    c_fin = torch.empty((ndir, B, U), dtype=torch.float32, device=x.device)
    h_fin = torch.empty((ndir, B, U), dtype=torch.float32, device=x.device)
    d = _lib.RecDesc()
    d.dtype, d.B, d.T, d.U, d.ndir = _lib.dtype_code(precision), B, T, U, ndir
    d.xproj, d.whh, d.lengths, d.out = xproj.data_ptr(), lw["whh"].data_ptr(), lengths.data_ptr(), out.data_ptr()
    d.out_batch_stride = out.stride(0)
    d.c_final, d.h_final = c_fin.data_ptr(), h_fin.data_ptr()
    d.whh_tc = lw["whh_tc"].data_ptr() if lw.get("whh_tc") is not None else None
    need = L.plas_rec_workspace_bytes(C.byref(d))
    ws = torch.empty((need,), dtype=torch.uint8, device=x.device)
    with _lib.stage("rec"):
        _lib.check(L.plas_bilstm_rec_fwd(C.byref(d), _lib.ptr(ws), need, _lib.stream_ptr()))
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


def listener(encoder_inputs, source_sequence_length, mode, hparams, weights):
    """las/model.py:104-142.  ``hparams`` is the flat dict (or the encoder view); ``weights`` a
    :class:`ListenerWeights`.  mode: 'train' | 'eval' | 'infer' (dropout must be 0 in 'train')."""
    if mode == "train" and float(hparams.get("dropout", 0.0)) > 0.0:
        raise NotImplementedError("input dropout in TRAIN mode (las/ops.py:14-18) is not built yet")
    return pyramidal_bilstm(encoder_inputs, source_sequence_length, mode, hparams, weights)
