"""Training step of the hot path on the GPU: ``las_model_fn(mode=TRAIN)`` without the Estimator glue.

Mirrors model_helper.py:165-227 (listener + speller(s) in TRAIN mode), :319-358 (sequence / sigmoid / CTC losses,
multitask sum), :403-417 (L2 regulariser, per-tensor ``clip_by_norm(grad, 2)``, Adam) and, for data-parallel runs,
the CrossShardOptimizer order of :405-406 (clip locally, then average the gradients across ranks, then apply).
Every stochastic op -- input dropout, scheduled sampling, bahdanau_monotonic's score noise, the periodic weight noise -- draws
from a counter-based hash keyed by (tensor, optimiser step): TF's RNG stream cannot be reproduced, so parity is checked against
the oracle replaying the same draws through the numpy mirrors below (``reference_masks``, ``reference_sampling``,
``reference_noise``, ``reference_weight_noise``).  Decoder variants: DESIGN.md section 8.

Every variable lives in ONE flat fp32 device buffer in the TF checkpoint layout; gradients and
the Adam moments mirror it, so the data-parallel exchange is a single all-reduce of ``state.grads`` and a trained
model exports to the reference's variable names unchanged.  All arithmetic runs in csrc/train_*.cu through the
C-ABI; torch only owns memory, streams and the process group.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .weights import variable_shapes

GRAD_NORM = 2.0  # model_helper.py:16


def _p(t, off_elems=0):
    return t.data_ptr() + 4 * off_elems


def gemm_ex(M, N, K, A, sam, sak, B, sbk, sbn, Cp, ldc, bias=None, beta=0.0, alpha=1.0, batch=1, ba=0, bb=0, bc=0, split_ws=None):
    """plas_gemm_f32_ex on raw device addresses (ints)."""
    d = _lib.GemmExDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.sam, d.sak = A, sam, sak
    d.B, d.sbk, d.sbn = B, sbk, sbn
    d.C, d.ldc, d.bias = Cp, ldc, bias
    d.alpha, d.beta = alpha, beta
    d.batch, d.batch_a, d.batch_b, d.batch_c = batch, ba, bb, bc
    if split_ws is not None:  # deterministic split-K scratch (weight gradients)
        d.split_ws, d.split_ws_bytes = split_ws.data_ptr(), split_ws.numel() * split_ws.element_size()
    _lib.check(_lib.lib().plas_gemm_f32_ex(C.byref(d), _lib.stream_ptr()))
    _lib.count_launches(1)


# --------------------------------------------------------------------------------------------------------------
# the big contractions of the listener on the tensor pipe: 3xTF32 (csrc/gemm_tf32.cu).  The operands are split (and, where
# the TF layout asks for it, transposed) into [hi | lo | hi] / [hi | hi | lo] copies first, so every product is the same TN
# tcgen05 GEMM over a 3x longer contraction axis.  Small or oddly aligned problems stay on the exact-fp32 SIMT kernel.
# --------------------------------------------------------------------------------------------------------------
TC_MIN_FLOPS = 1.0e8


def use_tc(M, N, K, *ptrs_lds):
    if os.environ.get("PLAS_TRAIN_GEMM") == "simt" or 2.0 * M * N * K < TC_MIN_FLOPS:
        return False
    return all(int(v) % 4 == 0 for v in ptrs_lds)  # 16-byte aligned bases (addresses are passed / 4) and leading dimensions


def split3(X, rows, cols, ld, pattern, transpose, device):
    """-> (tensor [rows or cols][3 * seg], seg): the [hi | lo | hi] (pattern 0) / [hi | hi | lo] (pattern 1) TF32 split of the
    fp32 matrix at device address X, optionally transposed; seg = contraction length rounded up to a whole k block."""
    inner = rows if transpose else cols
    seg = (inner + 31) // 32 * 32
    out = torch.empty((cols if transpose else rows, 3 * seg), dtype=torch.float32, device=device)
    _lib.check(_lib.lib().plas_split3_f32(C.c_void_p(X), rows, cols, ld, _lib.ptr(out), 3 * seg, seg, pattern, 1 if transpose else 0,
                                          _lib.stream_ptr()))
    _lib.count_launches(1)
    return out, seg


def gemm_tc(A3, B3, M, N, seg, Cp, ldc, bias=None, accumulate=False):
    """C[M][N] (+)= A3 . B3^T (+ bias) on pre-split operands."""
    L = _lib.lib()
    need = L.plas_gemm_tf32x3_scratch_bytes(M, N, 3 * seg)
    ws = torch.empty((need,), dtype=torch.uint8, device=A3.device) if need else None
    _lib.check(L.plas_gemm_tf32x3_tn(_lib.ptr(A3), M, 3 * seg, A3.shape[1], _lib.ptr(B3), N, B3.shape[1],
                                     C.c_void_p(bias) if bias else None, C.c_void_p(Cp), ldc, 1 if accumulate else 0,
                                     _lib.ptr(ws), need, _lib.stream_ptr()))
    _lib.count_launches(2 if need else 1)


def colsum(X, M, N, ld, out, accumulate=False):
    _lib.check(_lib.lib().plas_colsum_f32(C.c_void_p(X), M, N, ld, C.c_void_p(out), 1 if accumulate else 0, _lib.stream_ptr()))
    _lib.count_launches(1)


# --------------------------------------------------------------------------------------------------------------
# input dropout (DropoutWrapper(input_keep_prob = 1 - dropout) around every LSTMCell in TRAIN, las/ops.py:14-18)
# --------------------------------------------------------------------------------------------------------------
def seed_base(hp):
    """Base seed of the dropout masks, scheduled-sampling draws and attention noise of THIS replica: data-parallel ranks set
    hp['replica_id'] = rank so that their masks decorrelate, as independent workers' would.  The periodic weight noise
    (--add_noise) keeps the plain hp['dropout_seed']: it must be identical on every rank or the replicas would diverge."""
    return (int(hp.get("dropout_seed", 0)) + 1000003 * int(hp.get("replica_id", 0) or 0)) & 0xFFFFFFFF


def drop_seed(base, step, tensor_id):
    """32-bit seed of one dropped-out tensor at one optimiser step."""
    return int((int(base) * 0x9E3779B1 + int(step) * 0x85EBCA77 + int(tensor_id) * 0xC2B2AE3D + 0x165667B1) & 0xFFFFFFFF)


def hash_u24(n, seed):
    """numpy mirror of drop_hash() >> 8 in csrc/common.cuh for the indices 0..n-1: 24-bit uniform integers."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        sd = np.uint32(seed & 0xFFFFFFFF)
        x = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32) * np.uint32(0x9E3779B1) + (idx >> np.uint64(32)).astype(np.uint32) * np.uint32(0x85EBCA77) + sd
        x ^= x >> np.uint32(16); x *= np.uint32(0x85EBCA6B); x ^= x >> np.uint32(13); x *= np.uint32(0xC2B2AE35); x ^= x >> np.uint32(16)
        x += sd * np.uint32(0x27D4EB2F)
        x ^= x >> np.uint32(15); x *= np.uint32(0x2C1B3C6D); x ^= x >> np.uint32(12); x *= np.uint32(0x297A2D39); x ^= x >> np.uint32(15)
    return x >> np.uint32(8)


def dropout_mask(n, seed, keep_prob):
    """numpy mirror of drop_scale(): multipliers (1/keep or 0) of elements 0..n-1 of the tensor `seed`."""
    thresh = np.uint32(np.float32(keep_prob) * np.float32(16777216.0))
    return np.where(hash_u24(n, seed) < thresh, np.float32(1.0) / np.float32(keep_prob), np.float32(0.0)).astype(np.float32)


def reference_sampling(hp, step, B, S, V, speller_index=0):
    """Scheduled-sampling randomness of the device path at optimiser step ``step`` (test support): ``selected`` [B,S] bool (row b
    replaces its input of step t+1 by a sample drawn at step t) and the Gumbel noise [B,S,V] added to the logits of step t."""
    base = seed_base(hp)
    seed = drop_seed(base, step, SPELLER_TID + 10 * speller_index + 9)
    p = np.uint32(np.float32(hp["sampling_probability"]) * np.float32(16777216.0))
    selected = (hash_u24(B * S, seed) < p).reshape(B, S)
    u = (hash_u24(B * S * V, (seed + 1) & 0xFFFFFFFF).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    return selected, (-np.log(-np.log(u.astype(np.float64)))).reshape(B, S, V)


def reference_noise(hp, step, B, S, Tm, speller_index=0, scale=1.0):
    """bahdanau_monotonic's TRAIN-mode score noise (las/model.py:161-162, sigmoid_noise = 1) as the device path draws it at
    optimiser step ``step`` (test support): scale * N(0,1) [B,S,Tm], Box-Muller on two counter-hash uniforms (hash_normal in
    csrc/train_dec.cu)."""
    seed = drop_seed(seed_base(hp), step, SPELLER_TID + 10 * speller_index + 8)
    n = B * S * Tm
    u1 = (hash_u24(n, seed).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    u2 = (hash_u24(n, (seed + 1) & 0xFFFFFFFF).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    arg = (np.float32(6.283185307179586) * u2).astype(np.float64)
    return (scale * np.sqrt(-2.0 * np.log(u1.astype(np.float64))) * np.cos(arg)).reshape(B, S, Tm)


def dropout_(x, y, seed, keep_prob, step_dev=None):
    """y = x * mask(seed + step * 0x85EBCA77) / keep_prob on the device (plas_dropout_f32; ``step_dev`` = the TrainState's device
    step counter, None = 0); the same call on a gradient is the backward pass."""
    _lib.check(_lib.lib().plas_dropout_f32(_lib.ptr(x), _lib.ptr(y), x.numel(), seed, _lib.ptr(step_dev), keep_prob, _lib.stream_ptr()))
    _lib.count_launches(1)
    return y


LISTENER_TID = 1      # + 2*layer + direction
SPELLER_TID = 100     # + 10*speller index: +0 decoder inputs, +1 attention_{t-1}, +2+l output of decoder layer l


def reference_masks(hp, step, B, T, C, S, binf_count=0):
    """The multipliers the device path applies at optimiser step ``step``, as numpy arrays keyed like oracle/las_torch.py's
    ``masks`` argument (test support: lets the CPU oracle replay the stochastic op on identical masks)."""
    keep = 1.0 - float(hp.get("dropout", 0.0))
    base = seed_base(hp)
    U, V, Ud = hp["encoder_units"], hp["target_vocab_size"], hp["decoder_units"]
    out = {"listener": {}}
    ndir, pyr = (1 if hp.get("unidirectional") else 2), bool(hp.get("use_pyramidal", True))
    t, din = T, C
    for l in range(hp["encoder_layers"]):
        for dd in range(ndir):  # the mask covers the whole layer input; a stacked cell reads its own column slice of it
            out["listener"][(l, dd)] = dropout_mask(B * t * din, drop_seed(base, step, LISTENER_TID + 2 * l + dd), keep).reshape(B, t, din)
        if pyr:
            din = ndir * U if l == 0 else 2 * ndir * U
            if l != 0:
                t = (t + 1) // 2
        else:
            din = ndir * U
    D = int(hp.get("attention_layer_size") or din)  # width of the attention vector that is fed back (las/model.py:180-200)
    for si, (scope, E) in enumerate((("speller", int(hp.get("embedding_size") or 0) or V), ("speller_binf", binf_count))):
        if E <= 0:
            continue
        tid = SPELLER_TID + 10 * si
        m = {"x": dropout_mask(B * S * E, drop_seed(base, step, tid), keep).reshape(B, S, E),
             "att": dropout_mask(B * S * D, drop_seed(base, step, tid + 1), keep).reshape(B, S, D)}
        for l in range(hp["decoder_layers"] - 1):
            m[("h", l)] = dropout_mask(B * S * Ud, (drop_seed(base, step, tid + 1) + 1 + l) & 0xFFFFFFFF, keep).reshape(B, S, Ud)
        if hp.get("bottom_only"):  # AttentionMultiCell: cell l >= 1 drops its whole input [output below; old attention]
            for l in range(1, hp["decoder_layers"]):
                K = (D if l == 1 else Ud) + D
                m[("in", l)] = dropout_mask(B * S * K, (drop_seed(base, step, tid + 1) + l) & 0xFFFFFFFF, keep).reshape(B, S, K)
        out[scope] = m
    return out


class TrainState:
    """Flat parameter / gradient / Adam-moment buffers keyed by TF variable name."""

    def __init__(self, params, device="cuda"):
        _lib.require_cuda()
        self.names = list(params.keys())
        self.shapes = {k: tuple(np.shape(params[k])) for k in self.names}
        sizes = [int(np.prod(self.shapes[k])) if self.shapes[k] else 1 for k in self.names]
        # every tensor starts on a 16-byte boundary (vector loads in the kernels); padding stays zero
        offs, off = [], 0
        for n in sizes:
            offs.append(off)
            off += (n + 3) // 4 * 4
        self.total = off
        self.offsets_host = offs + [off]
        self.sizes = sizes
        flat = np.zeros((off,), np.float32)
        for k, o, n in zip(self.names, offs, sizes):
            flat[o:o + n] = np.asarray(params[k], np.float32).reshape(-1)
        self.params = torch.from_numpy(flat).to(device)
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.offsets = torch.tensor(self.offsets_host, dtype=torch.int64, device=device)
        self.norms = torch.zeros((len(self.names),), dtype=torch.float32, device=device)
        self.wsq = torch.zeros((len(self.names),), dtype=torch.float32, device=device)
        self.l2_scratch = torch.empty((_lib.lib().plas_grad_l2_norm_scratch_bytes(len(self.names)),), dtype=torch.uint8, device=device)
        self.index = {k: i for i, k in enumerate(self.names)}
        self.split_ws = torch.empty((32 << 20,), dtype=torch.uint8, device=device)  # split-K scratch of the dW GEMMs
        self.step_dev = torch.zeros((1,), dtype=torch.int32, device=device)  # read by the dropout kernels (graph-safe seeds)
        self._step = 0
        self._streams = []

    @property
    def step(self):
        """Number of optimiser steps applied so far (mirrored on the device for the dropout seeds)."""
        return self._step

    @step.setter
    def step(self, n):
        self._step = int(n)
        self.step_dev.fill_(int(n))

    def side_streams(self, n):
        while len(self._streams) < n:
            self._streams.append(torch.cuda.Stream(device=self.params.device))
        return self._streams[:n]

    def w(self, name, row=0):
        """device address of variable ``name`` (+ ``row`` rows of its last dimension)."""
        i = self.index[name]
        cols = self.shapes[name][-1] if self.shapes[name] else 1
        return _p(self.params, self.offsets_host[i] + row * cols)

    def view(self, name):
        """the parameter ``name`` as a tensor view of the flat buffer (gather source for embedding lookups)."""
        i = self.index[name]
        o = self.offsets_host[i]
        return self.params[o:o + int(np.prod(self.shapes[name]))].view(self.shapes[name])

    def g(self, name, row=0):
        i = self.index[name]
        cols = self.shapes[name][-1] if self.shapes[name] else 1
        return _p(self.grads, self.offsets_host[i] + row * cols)

    def export_params(self):
        """-> {tf_variable_name: float32 ndarray} (the exchange format of weights.py)."""
        flat = self.params.cpu().numpy()
        return {k: flat[o:o + n].reshape(self.shapes[k]).copy() for k, o, n in zip(self.names, self.offsets_host, self.sizes)}

    # -- TF checkpoint interchange (tf_checkpoint.py): variables under their TF names plus the slots tf.train.AdamOptimizer keeps
    def save_checkpoint(self, prefix):
        """Write model.ckpt-style files the reference's Estimator could warm-start from: every variable, its Adam moments as
        ``<name>/Adam`` and ``<name>/Adam_1``, ``beta1_power`` / ``beta2_power`` (= beta^(t+1) after t steps, as TF keeps them)
        and ``global_step``."""
        from . import tf_checkpoint
        out = self.export_params()
        m, v = self.m.cpu().numpy(), self.v.cpu().numpy()
        for k, o, n in zip(self.names, self.offsets_host, self.sizes):
            out[k + "/Adam"] = m[o:o + n].reshape(self.shapes[k]).copy()
            out[k + "/Adam_1"] = v[o:o + n].reshape(self.shapes[k]).copy()
        out["beta1_power"] = np.array(0.9 ** (self.step + 1), np.float32)
        out["beta2_power"] = np.array(0.999 ** (self.step + 1), np.float32)
        out["global_step"] = np.array(self.step, np.int64)
        tf_checkpoint.write_checkpoint(prefix, out)

    @classmethod
    def from_checkpoint(cls, prefix, names, device="cuda"):
        """Restore a TrainState for the variables ``names`` (ordered) from a TF bundle; missing Adam slots start at zero."""
        from . import tf_checkpoint
        ck = tf_checkpoint.read_checkpoint(prefix)
        st = cls({k: ck[k] for k in names}, device=device)
        m, v = np.zeros((st.total,), np.float32), np.zeros((st.total,), np.float32)
        for k, o, n in zip(st.names, st.offsets_host, st.sizes):
            if k + "/Adam" in ck:
                m[o:o + n] = ck[k + "/Adam"].reshape(-1)
                v[o:o + n] = ck[k + "/Adam_1"].reshape(-1)
        st.m.copy_(torch.from_numpy(m))
        st.v.copy_(torch.from_numpy(v))
        st.step = int(ck["global_step"]) if "global_step" in ck else 0
        return st

    def export_grads(self):
        flat = self.grads.cpu().numpy()
        return {k: flat[o:o + n].reshape(self.shapes[k]).copy() for k, o, n in zip(self.names, self.offsets_host, self.sizes)}


# --------------------------------------------------------------------------------------------------------------
# listener
# --------------------------------------------------------------------------------------------------------------
def _layer_names(l, hp):
    """TF variable scopes of the cells of listener layer l, one per direction (las/ops.py:23-46, las/model.py:111-142)."""
    uni = bool(hp["unidirectional"])
    if hp["use_pyramidal"]:
        dirs = ["rnn"] if uni else ["bidirectional_rnn/fw", "bidirectional_rnn/bw"]
        return [f"listener/bilstm_{l}/{d}/lstm_cell" for d in dirs]
    dirs = ["rnn"] if uni else ["bidirectional_rnn/fw", "bidirectional_rnn/bw"]
    return [f"listener/{d}/multi_rnn_cell/cell_{l}/lstm_cell" for d in dirs]


def _rec_desc(B, T, U, ndir, din, z, kernels, lengths, out, c_save, h_prev, dout=None):
    d = _lib.RecTrainDesc()
    d.B, d.T, d.U, d.ndir, d.din = B, T, U, ndir, din
    d.z = z.data_ptr()
    for i, k in enumerate(kernels):
        d.kernel[i] = k
    d.lengths = lengths.data_ptr()
    d.out = out.data_ptr() if out is not None else None
    d.out_batch_stride = (out if out is not None else dout).stride(0)
    d.c_save, d.h_prev = c_save.data_ptr(), h_prev.data_ptr()
    d.dout = dout.data_ptr() if dout is not None else None
    return d


def listener_train_fwd(x, lengths, st, hp):
    """listener (las/model.py:104-142) forward keeping what BPTT needs: pyramidal_bilstm (las/ops.py:68-87) or the stacked
    MultiRNNCell listener, bidirectional or unidirectional.  x [B,T,C] f32 -> (enc_out, enc_len, tape).
    With hp['dropout'] > 0 every cell sees its own dropped-out copy of its input (one DropoutWrapper per cell).
    In the stacked (non-pyramidal) listener each direction is its own stack: from layer 1 on, direction d reads only the U
    columns of its own previous output -- a strided view for the GEMMs, no copy."""
    L = _lib.lib()
    U, ndir, pyr = hp["encoder_units"], (1 if hp["unidirectional"] else 2), bool(hp["use_pyramidal"])
    B = x.shape[0]
    lengths = lengths.to(device=x.device, dtype=torch.int32).contiguous()
    x = x.to(torch.float32).contiguous()
    keep = 1.0 - float(hp.get("dropout", 0.0))
    base = seed_base(hp)
    tape = []
    for l in range(hp["encoder_layers"]):
        T, width = x.shape[1], x.shape[2]
        own = (not pyr) and l > 0           # direction d reads columns [d*U, (d+1)*U) of x
        din = U if own else width           # input depth of one cell
        names = _layer_names(l, hp)
        z = torch.empty((B, T, ndir, 4 * U), dtype=torch.float32, device=x.device)
        seeds = [drop_seed(base, 0, LISTENER_TID + 2 * l + dd) for dd in range(ndir)]  # + step * DROP_STEP_MUL on the device
        xs = [dropout_(x, torch.empty_like(x), seeds[dd], keep, st.step_dev) for dd in range(ndir)] if keep < 1.0 else [x] * ndir
        with _lib.stage("train_inproj"):
            for dd, nm in enumerate(names):
                xa = _p(xs[dd], dd * U if own else 0)
                if use_tc(B * T, 4 * U, din, xa // 4, width, st.w(nm + "/bias") // 4):  # x W + b = [x split] . [W^T split]^T
                    a3, seg = split3(xa, B * T, din, width, 0, False, x.device)
                    b3, _ = split3(st.w(nm + "/kernel"), din, 4 * U, 4 * U, 1, True, x.device)
                    gemm_tc(a3, b3, B * T, 4 * U, seg, _p(z, dd * 4 * U), ndir * 4 * U, bias=st.w(nm + "/bias"))
                else:
                    gemm_ex(B * T, 4 * U, din, xa, width, 1, st.w(nm + "/kernel"), 4 * U, 1,
                            _p(z, dd * 4 * U), ndir * 4 * U, bias=st.w(nm + "/bias"))
        stack = pyr and l != 0
        t_alloc = T + (T % 2) if stack else T
        out = torch.zeros((B, t_alloc, ndir * U), dtype=torch.float32, device=x.device)
        c_save = torch.zeros((B, T, ndir * U), dtype=torch.float32, device=x.device)
        h_prev = torch.zeros((B, T, ndir * U), dtype=torch.float32, device=x.device)
        d = _rec_desc(B, T, U, ndir, din, z, [st.w(nm + "/kernel") for nm in names], lengths, out, c_save, h_prev)
        c_fin = torch.empty((ndir, B, U), dtype=torch.float32, device=x.device)  # encoder_state of the layer (pass_hidden_state)
        h_fin = torch.empty((ndir, B, U), dtype=torch.float32, device=x.device)
        d.c_final, d.h_final = c_fin.data_ptr(), h_fin.data_ptr()
        need = L.plas_rec_train_workspace_bytes(C.byref(d))
        ws = torch.empty((need,), dtype=torch.uint8, device=x.device)
        with _lib.stage("train_rec_fwd"):
            _lib.check(L.plas_bilstm_rec_train_fwd(C.byref(d), _lib.ptr(ws), need, _lib.stream_ptr()))
        _lib.count_launches(1)
        tape.append(dict(xs=xs, seeds=seeds, keep=keep, z=z, c_save=c_save, h_prev=h_prev, lengths=lengths, T=T, din=din,
                         width=width, own=own, t_alloc=t_alloc, ws=ws, c_fin=c_fin, h_fin=h_fin))
        if stack:  # pyramidal_stack: free view + ceil-halved lengths (las/ops.py:49-65)
            out = out.view(B, t_alloc // 2, 2 * ndir * U)
            lengths = torch.div(lengths, 2, rounding_mode="floor") + lengths % 2
        x = out
    return x, lengths, tape


def listener_train_bwd(d_enc, tape, st, hp, d_final=None):
    """Gradient of the listener: per layer BPTT (plas_bilstm_rec_train_bwd) then dW = [x;h]^T dz, db, dx = dz W_x^T.
    ``d_final`` = (dc [ndir,B,U], dh [ndir,B,U]): gradients wrt the LAST layer's final states (pass_hidden_state)."""
    L = _lib.lib()
    U, ndir = hp["encoder_units"], (1 if hp["unidirectional"] else 2)
    B = d_enc.shape[0]
    dout = d_enc
    side = st.side_streams(3)[2]
    for l in range(hp["encoder_layers"] - 1, -1, -1):
        tp = tape[l]
        T, din, width, own = tp["T"], tp["din"], tp["width"], tp["own"]
        names = _layer_names(l, hp)
        dout = dout.reshape(B, tp["t_alloc"], ndir * U)
        d = _rec_desc(B, T, U, ndir, din, tp["z"], [st.w(nm + "/kernel") for nm in names], tp["lengths"], None, tp["c_save"],
                      tp["h_prev"], dout=dout)
        if d_final is not None and l == hp["encoder_layers"] - 1:
            d.dc_final, d.dh_final = d_final[0].data_ptr(), d_final[1].data_ptr()
        need = tp["ws"].numel()
        with _lib.stage("train_rec_bwd"):
            _lib.check(L.plas_bilstm_rec_train_bwd(C.byref(d), _lib.ptr(tp["ws"]), need, _lib.stream_ptr()))
        _lib.count_launches(1)
        z, xs, hp_, keep = tp["z"], tp["xs"], tp["h_prev"], tp["keep"]
        M = B * T
        main = torch.cuda.current_stream()
        if l > 0:  # input gradient first: the next recurrence depends on it
            dx = torch.empty((B, T, width), dtype=torch.float32, device=z.device)
            with _lib.stage("train_dgrad"):
                tc_d = use_tc(M, din, 4 * U, width, U)
                for dd, nm in enumerate(names):
                    zp, wk = _p(z, dd * 4 * U), st.w(nm + "/kernel")
                    if tc_d:  # dz W^T = [dz split] . [W split]^T: W [din][4U] already has the contraction index contiguous
                        dz3, segz = split3(zp, M, 4 * U, ndir * 4 * U, 0, False, z.device)
                        w3, _ = split3(wk, din, 4 * U, 4 * U, 1, False, z.device)
                    if keep < 1.0 and tc_d:
                        dxd = dx if dd == 0 else torch.empty_like(dx)
                        if own:
                            dxd.zero_()
                        gemm_tc(dz3, w3, M, din, segz, _p(dxd, dd * U if own else 0), width)
                        dropout_(dxd, dxd, tp["seeds"][dd], keep, st.step_dev)
                        if dd > 0:
                            _lib.check(L.plas_axpy_f32(_lib.ptr(dx), _lib.ptr(dxd), dx.numel(), 1.0, _lib.stream_ptr()))
                            _lib.count_launches(1)
                    elif tc_d:
                        gemm_tc(dz3, w3, M, din, segz, _p(dx, dd * U if own else 0), width, accumulate=(dd > 0 and not own))
                    elif keep < 1.0:  # each cell saw its own mask: dx = sum_d mask_d * (dz_d W_d^T)
                        dxd = dx if dd == 0 else torch.empty_like(dx)
                        if own:
                            dxd.zero_()
                        gemm_ex(M, din, 4 * U, zp, ndir * 4 * U, 1, wk, 1, 4 * U, _p(dxd, dd * U if own else 0), width)
                        dropout_(dxd, dxd, tp["seeds"][dd], keep, st.step_dev)
                        if dd > 0:
                            _lib.check(L.plas_axpy_f32(_lib.ptr(dx), _lib.ptr(dxd), dx.numel(), 1.0, _lib.stream_ptr()))
                            _lib.count_launches(1)
                    elif own:  # the directions' inputs are disjoint column slices
                        gemm_ex(M, din, 4 * U, zp, ndir * 4 * U, 1, wk, 1, 4 * U, _p(dx, dd * U), width)
                    else:
                        gemm_ex(M, din, 4 * U, zp, ndir * 4 * U, 1, wk, 1, 4 * U, dx.data_ptr(), width, beta=0.0 if dd == 0 else 1.0)
            dout = dx
        # weight gradients on a side stream: they only need dz, and overlap the (latency-bound) recurrence of the layer below
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            with _lib.stage("train_wgrad"):
                for dd, nm in enumerate(names):
                    zp = _p(z, dd * 4 * U)
                    xa = _p(xs[dd], dd * U if own else 0)
                    if use_tc(din + U, 4 * U, M, xa // 4, width, st.g(nm + "/kernel") // 4):
                        # X^T dZ = [X^T split] . [dZ^T split]^T: both operands transposed into the TN form; one CTA walks the whole
                        # contraction of its output tile, so the sum order is fixed (deterministic without split-K scratch)
                        dzt3, segm = split3(zp, M, 4 * U, ndir * 4 * U, 1, True, z.device)
                        xt3, _ = split3(xa, M, din, width, 0, True, z.device)
                        gemm_tc(xt3, dzt3, din, 4 * U, segm, st.g(nm + "/kernel"), 4 * U)
                        ht3, _ = split3(_p(hp_, dd * U), M, U, ndir * U, 0, True, z.device)
                        gemm_tc(ht3, dzt3, U, 4 * U, segm, st.g(nm + "/kernel", din), 4 * U)
                    else:
                        gemm_ex(din, 4 * U, M, xa, 1, width, zp, ndir * 4 * U, 1, st.g(nm + "/kernel"), 4 * U, split_ws=st.split_ws)
                        gemm_ex(U, 4 * U, M, _p(hp_, dd * U), 1, ndir * U, zp, ndir * 4 * U, 1, st.g(nm + "/kernel", din), 4 * U,
                                split_ws=st.split_ws)
                    colsum(zp, M, 4 * U, ndir * 4 * U, st.g(nm + "/bias"))
    done = torch.cuda.Event()
    done.record(side)
    torch.cuda.current_stream().wait_event(done)


# --------------------------------------------------------------------------------------------------------------
# speller
# --------------------------------------------------------------------------------------------------------------
class SpellerTrain:
    """One teacher-forced speller (scope 'speller' or 'speller_binf') bound to a TrainState."""

    def __init__(self, st, hp, scope, E, n_out, index=0, binf=None, table=None):
        """``binf`` (device tensor [n, V], binf2phone): --binf_projection wiring of this speller (las/model.py:251-257): the
        projection is the constant transform_binf_to_phones map [M; 1 - M] on the 2n-wide attention vectors, which ``forward``
        keeps in ``self.att_vec`` for the log-probability regulariser."""
        self.st, self.hp, self.scope, self.E, self.n_out = st, hp, scope, E, n_out
        self.proj_const = self.att_vec = self.dproj = None
        if binf is not None:
            M = binf.to(torch.float32)
            if hp.get("binf_trainable"):  # the matrix is the variable 'binf2phone' (model_helper.py:183): keep d([M; 1 - M])
                self.dproj = torch.empty((2 * M.shape[0], M.shape[1]), dtype=torch.float32, device=M.device)
            self.proj_const = torch.cat([M, 1.0 - M], 0).contiguous()  # [2n, V]
            self.zero_bias = torch.zeros((M.shape[1],), dtype=torch.float32, device=M.device)
            assert int(hp.get("attention_layer_size") or 0) == self.proj_const.shape[0] and n_out == M.shape[1]
        self.keep = 1.0 - float(hp.get("dropout", 0.0))
        self.tid = SPELLER_TID + 10 * index
        self.base = seed_base(hp)
        self.init = self.d_init = None
        # scheduled sampling (las/model.py:279-288): the phone speller draws ids from its own logits; the binary-feature
        # speller's ScheduledSigmoidHelper path of the reference is shape-inconsistent and is not built
        # ``table`` [n_out, E]: embedding_fn when the inputs are not one-hot (target_embedding, or the binary-feature columns of
        # --binf_projection) -- a sampled id then feeds its row, and ``self.fed_ids`` records the ids actually fed
        self.sample_prob = float(hp.get("sampling_probability", 0.0))
        self.table, self.fed_ids = table, None
        if self.sample_prob > 0.0 and E != n_out and table is None:
            raise NotImplementedError("training path: scheduled sampling of the binary-feature speller (ScheduledSigmoidHelper) is not built; "
                                      "set sampling_probability=0")
        if hp["attention_type"] not in _lib.ATT_CODES:
            raise NotImplementedError(f"training path: attention_type={hp['attention_type']}")
        # bahdanau_monotonic in TRAIN mode: sigmoid_noise = 1.0 (las/model.py:161-162); tests may set 0 for a noise-free check
        self.sigmoid_noise = 1.0 if hp["attention_type"] == "bahdanau_monotonic" else 0.0
        if hp.get("embedding_size") and hp.get("binary_outputs"):
            # the reference's embedding_fn looks float feature vectors up in target_embedding there (las/model.py:229-237): ill-formed
            raise NotImplementedError("training path: --embedding_size with binary_outputs is not built")
        self.dx_in = None
        self.att_layer = int(hp.get("attention_layer_size") or 0)
        self.bottom = bool(hp.get("bottom_only"))
        self.pass_state = self.bottom and bool(hp.get("pass_hidden_state"))  # las/model.py:260 needs both flags

    def _desc(self, memory, mem_len, x_in, logits, dlogits=None, dmemory=None):
        st, hp, sc = self.st, self.hp, self.scope
        B, Tm, D = memory.shape
        d = _lib.DecTrainDesc()
        d.B, d.S, d.Tm, d.D, d.Ud, d.E, d.n_out = B, x_in.shape[1], Tm, D, hp["decoder_units"], self.E, self.n_out
        d.n_layers = hp["decoder_layers"]
        d.attention_type = _lib.ATT_CODES[hp["attention_type"]]
        d.dmemory_accumulate = 1
        d.keep_prob = self.keep
        d.drop_seed = drop_seed(self.base, 0, self.tid + 1)  # + step * DROP_STEP_MUL on the device
        d.drop_step = st.step_dev.data_ptr()
        d.sample_prob = self.sample_prob
        if self.sample_prob > 0.0 and self.fed_ids is not None:  # the ids actually fed (teacher or drawn), step by step
            d.sample_fed_ids = self.fed_ids.data_ptr()
        if self.sample_prob > 0.0 and self.table is not None:
            d.sample_table = self.table.data_ptr()
        d.sample_seed = drop_seed(self.base, 0, self.tid + 9)
        d.xdrop_seed = drop_seed(self.base, 0, self.tid)
        d.x_in_rw = x_in.data_ptr()
        d.bottom_only = 1 if self.bottom else 0
        pre = f"{sc}/decoder/multi_rnn_cell/cell_0_attention/attention_wrapper" if self.bottom else f"{sc}/decoder/attention_wrapper"
        for k in range(d.n_layers):
            if self.bottom:  # AttentionMultiCell scopes (las/model.py:43-54)
                nm = f"{pre}/lstm_cell" if k == 0 else f"{sc}/decoder/multi_rnn_cell/cell_{k}/lstm_cell"
            else:
                nm = f"{pre}/multi_rnn_cell/cell_{k}/lstm_cell"
            d.kernel[k], d.bias[k] = st.w(nm + "/kernel"), st.w(nm + "/bias")
            d.dkernel[k], d.dbias[k] = st.g(nm + "/kernel"), st.g(nm + "/bias")
        d.w_mem, d.dw_mem = st.w(f"{sc}/memory_layer/kernel"), st.g(f"{sc}/memory_layer/kernel")
        if self.att_layer:  # AttentionWrapper's attention_layer (las/model.py:180-200)
            al = f"{pre}/attention_layer/kernel"
            d.att_layer, d.w_att_layer, d.dw_att_layer = self.att_layer, st.w(al), st.g(al)
        at = hp["attention_type"]
        if at in ("bahdanau", "bahdanau_monotonic"):
            q, v = f"{pre}/{at}_attention/query_layer/kernel", f"{pre}/{at}_attention/attention_v"
            d.w_query, d.v_att, d.dw_query, d.dv_att = st.w(q), st.w(v), st.g(q), st.g(v)
        elif at == "custom":  # CustomAttention's own query layer (las/model.py:92-93)
            q = f"{pre}/query_layer/kernel"
            d.w_query, d.dw_query = st.w(q), st.g(q)
        if at.endswith("_monotonic"):
            sb = f"{pre}/{at}_attention/attention_score_bias"
            d.score_bias, d.dscore_bias = st.w(sb), st.g(sb)
            d.sigmoid_noise, d.noise_seed = self.sigmoid_noise, drop_seed(self.base, 0, self.tid + 8)
        if self.proj_const is not None:  # constant projection: the Dense variables exist (checkpoint layout) but are never used
            d.w_proj, d.b_proj, d.db_proj = self.proj_const.data_ptr(), self.zero_bias.data_ptr(), None
            d.dw_proj = self.dproj.data_ptr() if self.dproj is not None else None
            d.att_out = self.att_vec.data_ptr()
            d.datt_extra = self.datt_extra.data_ptr() if self.datt_extra is not None else None
        else:
            pk, pb = f"{sc}/decoder/projection_layer/kernel", f"{sc}/decoder/projection_layer/bias"
            d.w_proj, d.b_proj, d.dw_proj, d.db_proj = st.w(pk), st.w(pb), st.g(pk), st.g(pb)
        d.memory, d.mem_len, d.x_in = memory.data_ptr(), mem_len.data_ptr(), x_in.data_ptr()
        d.logits = logits.data_ptr()
        d.dlogits = dlogits.data_ptr() if dlogits is not None else None
        d.dx_in = self.dx_in.data_ptr() if (self.dx_in is not None and dlogits is not None) else None
        d.dmemory = dmemory.data_ptr() if dmemory is not None else None
        if self.init is not None:
            for l, (c0, h0) in enumerate(self.init):
                d.c_init[l], d.h_init[l] = c0.data_ptr(), h0.data_ptr()
                if self.d_init is not None:
                    d.dc_init[l], d.dh_init[l] = self.d_init[l][0].data_ptr(), self.d_init[l][1].data_ptr()
        return d

    def forward(self, memory, mem_len, x_in, initial_state=None, ids=None):
        """memory [B,Tm,D] (zero past mem_len), x_in [B,S,E] -> logits [B,S,n_out]; keeps the tape on self.
        ``initial_state``: [(c, h)] per decoder cell (pass_hidden_state: the listener's final fw / bw states).
        ``ids`` [B,S]: the teacher ids behind ``x_in`` (needed with scheduled sampling through an embedding ``table``)."""
        L = _lib.lib()
        if self.sample_prob > 0.0 and self.table is not None:
            self.table = self.table.to(torch.float32).contiguous()
        if self.sample_prob > 0.0 and (self.table is not None or ids is not None):
            self.fed_ids = ids.to(torch.int32).contiguous().clone()
        self.init, self.d_init = None, None
        if self.pass_state and initial_state is not None:
            self.init = [(c.contiguous(), h.contiguous()) for c, h in initial_state[:self.hp["decoder_layers"]]]
            assert self.init[0][0].shape[1] == self.hp["decoder_units"], "pass_hidden_state needs encoder_units == decoder_units"
        B, S = x_in.shape[0], x_in.shape[1]
        x_in = x_in.contiguous()
        if self.keep < 1.0:  # the cell input [x_t; attention_{t-1}] is dropped out; x_t here, attention inside the kernels
            x_in = dropout_(x_in, torch.empty_like(x_in), drop_seed(self.base, 0, self.tid), self.keep, self.st.step_dev)
        self.memory, self.mem_len, self.x_in = memory.contiguous(), mem_len, x_in
        self.logits = torch.empty((B, S, self.n_out), dtype=torch.float32, device=memory.device)
        self.datt_extra = None
        if self.proj_const is not None:
            self.att_vec = torch.empty((B, S, self.proj_const.shape[0]), dtype=torch.float32, device=memory.device)
        d = self._desc(self.memory, self.mem_len, self.x_in, self.logits)
        need = L.plas_dec_train_workspace_bytes(C.byref(d))
        self.ws = torch.empty((need,), dtype=torch.uint8, device=memory.device)
        with _lib.stage("train_dec_fwd"):
            _lib.check(L.plas_decoder_train_fwd(C.byref(d), _lib.ptr(self.ws), need, _lib.stream_ptr()))
        _lib.count_launches(3 + S * (self.hp["decoder_layers"] + 1))
        return self.logits

    def backward(self, dlogits, d_enc, datt_extra=None, want_dx=False):
        """Accumulates into d_enc [B,Tm,D] and writes this speller's weight gradients into the TrainState; with
        pass_hidden_state the gradients wrt the initial states land in ``self.d_init`` [(dc, dh)] per seeded cell.
        ``datt_extra`` [B,S,A] (--binf_projection): gradient that reaches the attention vectors besides the projection's."""
        L = _lib.lib()
        self.datt_extra = None if datt_extra is None else datt_extra.contiguous()
        # want_dx: also the gradient wrt the decoder inputs (through their dropout mask), left in self.dx_in [B,S,E]
        self.dx_in = torch.empty_like(self.x_in) if want_dx else None
        if self.init is not None:
            self.d_init = [(torch.empty_like(c), torch.empty_like(h)) for c, h in self.init]
        d = self._desc(self.memory, self.mem_len, self.x_in, self.logits, dlogits.contiguous(), d_enc)
        with _lib.stage("train_dec_bwd"):
            _lib.check(L.plas_decoder_train_bwd(C.byref(d), _lib.ptr(self.ws), self.ws.numel(), _lib.stream_ptr()))
        _lib.count_launches(12 + self.x_in.shape[1] * (1 + 2 * self.hp["decoder_layers"]))
        if want_dx and self.keep < 1.0:
            dropout_(self.dx_in, self.dx_in, drop_seed(self.base, 0, self.tid), self.keep, self.st.step_dev)


# --------------------------------------------------------------------------------------------------------------
# losses with gradients
# --------------------------------------------------------------------------------------------------------------
def _seq_mask(lengths, maxlen):
    ar = torch.arange(maxlen, device=lengths.device)
    return (ar[None, :] < lengths[:, None].to(ar.dtype)).to(torch.float32).contiguous()


def seq_ce_grad(logits, targets, weights, gscale=1.0):
    B, S, V = logits.shape
    ce = torch.empty((B * S,), dtype=torch.float32, device=logits.device)
    out3 = torch.empty((3,), dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits)
    _lib.check(_lib.lib().plas_seq_ce_grad(_lib.ptr(logits), _lib.ptr(targets), _lib.ptr(weights), B * S, V, gscale, _lib.ptr(ce),
                                           _lib.ptr(out3), _lib.ptr(dl), _lib.stream_ptr()))
    _lib.count_launches(3)
    return out3[0], dl


def log_probs_reg_grad(att, weight=1.0):
    """compute_log_probs_loss (model_helper.py:132-146) of the attention vectors [B,S,2n] -> (weight * loss, d/d att)."""
    B, S, A = att.shape
    rows = torch.empty((B * S,), dtype=torch.float32, device=att.device)
    out3 = torch.empty((3,), dtype=torch.float32, device=att.device)
    datt = torch.empty_like(att)
    _lib.check(_lib.lib().plas_log_probs_reg_grad(_lib.ptr(att), B * S, A // 2, float(weight), _lib.ptr(rows), _lib.ptr(out3),
                                                  _lib.ptr(datt), _lib.stream_ptr()))
    _lib.count_launches(2)
    return out3[0], datt


def sigmoid_ce_grad(logits, labels, weights, gscale=1.0):
    B, S, n = logits.shape
    ce = torch.empty((B * S,), dtype=torch.float32, device=logits.device)
    out3 = torch.empty((3,), dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits)
    _lib.check(_lib.lib().plas_sigmoid_ce_grad(_lib.ptr(logits), _lib.ptr(labels), _lib.ptr(weights), B * S, n, gscale, _lib.ptr(ce),
                                               _lib.ptr(out3), _lib.ptr(dl), _lib.stream_ptr()))
    _lib.count_launches(3)
    return out3[0], dl


def ctc_grad(logits, labels, label_length, logit_length, gscale=1.0, blank=0):
    L = _lib.lib()
    B, T, Cn = logits.shape
    Lmax = labels.shape[1]
    loss = torch.empty((B,), dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits)
    need = L.plas_ctc_grad_workspace_bytes(B, T, Lmax)
    ws = torch.empty((max(need, 4),), dtype=torch.uint8, device=logits.device)
    _lib.check(L.plas_ctc_grad(_lib.ptr(logits), _lib.ptr(labels), _lib.ptr(label_length), _lib.ptr(logit_length), B, T, Cn, Lmax,
                               blank, gscale, _lib.ptr(loss), _lib.ptr(dl), _lib.ptr(ws), need, _lib.stream_ptr()))
    _lib.count_launches(1)
    return loss, dl


# --------------------------------------------------------------------------------------------------------------
# the step
# --------------------------------------------------------------------------------------------------------------
def forward_backward(features, labels, st, hp, binf=None):
    """Forward + backward of las_model_fn(TRAIN): fills ``st.grads`` (raw, before L2 / clipping) and returns the loss
    parts as device scalars {ce, ce_binf, ctc, audio_loss}.  ``binf`` [n, V] float tensor (binf2phone, model_helper.py:179-186)
    enables the multitask binary-feature speller when hp['binary_outputs']."""
    x, lens = features["encoder_inputs"], features["source_sequence_length"]
    dev = x.device
    tin = labels["targets_inputs"].to(device=dev, dtype=torch.int64)
    tout = labels["targets_outputs"].to(device=dev, dtype=torch.int32).contiguous()
    tlen = labels["target_sequence_length"].to(device=dev, dtype=torch.int32).contiguous()
    V = hp["target_vocab_size"]
    S = tin.shape[1]
    st.grads.zero_()
    enc_out, enc_len, tape = listener_train_fwd(x, lens, st, hp)
    enc_out = enc_out.contiguous()
    B, Tm, D = enc_out.shape
    w = _seq_mask(tlen, S)
    d_enc = torch.zeros_like(enc_out)
    # encoder_state (las/ops.py:68-87): final (c, h) of the last listener layer per direction -- seeds the decoder cells with
    # --pass_hidden_state --bottom_only (las/model.py:259-267)
    enc_state = [(tape[-1]["c_fin"][dd], tape[-1]["h_fin"][dd]) for dd in range(tape[-1]["c_fin"].shape[0])]
    parts = {}
    # the heads only share the encoder outputs: each speller (forward, loss, backward) runs on its own stream while the
    # CTC head runs on the caller's, each accumulating into its own encoder-output gradient buffer
    jobs = []
    emb = bool(hp.get("embedding_size"))
    onehot = torch.nn.functional.one_hot(tin, V).to(torch.float32)
    if not hp.get("binary_outputs") or hp.get("multitask"):
        # embedding_fn (las/model.py:228-246): one-hot ids, or rows of speller/target_embedding [V, E]
        x_sp = st.view("speller/target_embedding")[tin].contiguous() if emb else onehot
        jobs.append(("speller", "ce", V, x_sp, None))
    proj = bool(hp.get("binf_projection"))
    trainable_binf = proj and bool(hp.get("binf_trainable"))
    if trainable_binf:  # binf2phone is a variable (model_helper.py:181-186): the live parameter replaces the constant
        binf = st.view("binf2phone")
    if hp.get("binary_outputs"):
        bt = binf.to(device=dev, dtype=torch.float32).t().contiguous()  # [V, n]
        if proj:  # model_helper.py:221-227: phone ids in (embedded as binary-feature columns), phone logits out
            jobs.append(("speller_binf", "ce_binf", V, bt[tin], None))
        else:
            jobs.append(("speller_binf", "ce_binf", bt.shape[1], bt[tin], bt[tout.long()].contiguous()))
    main = torch.cuda.current_stream()
    d_enc_heads = [torch.zeros_like(enc_out) for _ in jobs]  # zero-filled on the caller's stream BEFORE the fork event
    ready = torch.cuda.Event()
    ready.record(main)
    side = st.side_streams(len(jobs))
    pending = []
    for (scope, key, n_out, x_in, lab), stream, d_enc_j in zip(jobs, side, d_enc_heads):
        with torch.cuda.stream(stream):
            stream.wait_event(ready)
            binf_proj = proj and scope == "speller_binf"
            table = st.view("speller/target_embedding") if (emb and scope == "speller") else (bt if binf_proj else None)
            sp = SpellerTrain(st, hp, scope, x_in.shape[2], n_out, index=0 if scope == "speller" else 1,
                              binf=binf.to(device=dev) if binf_proj else None, table=table)
            logits = sp.forward(enc_out, enc_len, x_in, initial_state=enc_state, ids=tin)
            # the ids that were actually fed (scheduled sampling replaces some): the input-side gradients scatter by them
            fed_onehot = onehot if sp.fed_ids is None else torch.nn.functional.one_hot(sp.fed_ids.long(), V).to(torch.float32)
            datt_extra = None
            if lab is None:
                parts[key], dl = seq_ce_grad(logits, tout, w)
                parts["logits" if scope == "speller" else "logits_binf"] = logits
                if binf_proj:  # + compute_log_probs_loss(raw attention) * binf_projection_reg_weight (model_helper.py:326-331)
                    parts["log_probs_reg"], datt_extra = log_probs_reg_grad(sp.att_vec, float(hp.get("binf_projection_reg_weight", 1.0)))
            else:
                parts[key], dl = sigmoid_ce_grad(logits, lab, w)
                parts["logits_binf"] = logits
            want_dx = (emb and scope == "speller") or (binf_proj and trainable_binf)
            sp.backward(dl, d_enc_j, datt_extra, want_dx=want_dx)
            if binf_proj and trainable_binf:
                # d(binf2phone) = dW[:n] - dW[n:] (projection [M; 1 - M]) + (OneHot^T dX)^T (the inputs are columns of M)
                n_b, gM = bt.shape[1], st.g("binf2phone")
                for rows, sign in ((sp.dproj[:n_b], 1.0), (sp.dproj[n_b:], -1.0)):
                    _lib.check(_lib.lib().plas_axpy_f32(C.c_void_p(gM), _lib.ptr(rows), rows.numel(), sign, _lib.stream_ptr()))
                    _lib.count_launches(1)
                gemm_ex(n_b, V, B * S, sp.dx_in.data_ptr(), 1, n_b, fed_onehot.data_ptr(), V, 1, gM, V, beta=1.0)
                continue_embedding = False
            else:
                continue_embedding = True
            if want_dx and continue_embedding:  # d(target_embedding)[v] = sum of dX over the positions that fed phone v: OneHot^T dX
                E = x_in.shape[2]
                gemm_ex(V, E, B * S, fed_onehot.data_ptr(), 1, V, sp.dx_in.data_ptr(), E, 1, st.g("speller/target_embedding"), E)
            done = torch.cuda.Event()
            done.record(stream)
        pending.append((d_enc_j, done, sp))
    total = None
    if hp.get("ctc_weight", -1.0) > 0:  # model_helper.py:347-358
        cw = float(hp["ctc_weight"])
        Cn = V + 1
        cl = torch.empty((B, Tm, Cn), dtype=torch.float32, device=dev)
        gemm_ex(B * Tm, Cn, D, enc_out.data_ptr(), D, 1, st.w("ctc_logits/kernel"), Cn, 1, cl.data_ptr(), Cn, bias=st.w("ctc_logits/bias"))
        per_utt, dcl = ctc_grad(cl, tout, tlen, enc_len, gscale=cw / B)
        parts["ctc"] = per_utt.mean()
        total = parts["ctc"] * cw
        gemm_ex(D, Cn, B * Tm, enc_out.data_ptr(), 1, D, dcl.data_ptr(), Cn, 1, st.g("ctc_logits/kernel"), Cn, split_ws=st.split_ws)
        colsum(dcl.data_ptr(), B * Tm, Cn, Cn, st.g("ctc_logits/bias"))
        gemm_ex(B * Tm, D, Cn, dcl.data_ptr(), Cn, 1, st.w("ctc_logits/kernel"), 1, Cn, d_enc.data_ptr(), D, beta=1.0)
    for (d_enc_j, done, _), (_, key, _, _, _) in zip(pending, jobs):
        main.wait_event(done)
        _lib.check(_lib.lib().plas_axpy_f32(_lib.ptr(d_enc), _lib.ptr(d_enc_j), d_enc.numel(), 1.0, _lib.stream_ptr()))
        _lib.count_launches(1)
        total = parts[key] if total is None else total + parts[key]
    if "log_probs_reg" in parts:
        total = total + parts["log_probs_reg"]
    d_final = None
    for _, _, sp in pending:  # gradients wrt the listener's final states from the seeded decoder cells
        if sp.d_init is not None:
            if d_final is None:
                d_final = (torch.zeros_like(tape[-1]["c_fin"]), torch.zeros_like(tape[-1]["h_fin"]))
            for l, (dc0, dh0) in enumerate(sp.d_init):
                for buf, g in ((d_final[0][l], dc0), (d_final[1][l], dh0)):
                    _lib.check(_lib.lib().plas_axpy_f32(_lib.ptr(buf), _lib.ptr(g), g.numel(), 1.0, _lib.stream_ptr()))
                    _lib.count_launches(1)
    listener_train_bwd(d_enc, tape, st, hp, d_final)
    parts["audio_loss"] = total
    parts["encoder_out"] = enc_out
    return parts


def regularise_and_clip(st, hp, world_size=1):
    """g += l2*w (model_helper.py:411-413); per-tensor clip_by_norm(g, 2) (:416); pre-scale by 1/world for the cross-rank mean."""
    L = _lib.lib()
    n = len(st.names)
    _lib.check(L.plas_grad_l2_norm(_lib.ptr(st.params), _lib.ptr(st.grads), _lib.ptr(st.offsets), n, float(hp.get("l2_reg_scale", 0.0)),
                                   _lib.ptr(st.norms), _lib.ptr(st.wsq), _lib.ptr(st.l2_scratch), st.l2_scratch.numel(), _lib.stream_ptr()))
    _lib.check(L.plas_clip_scale(_lib.ptr(st.grads), _lib.ptr(st.offsets), n, _lib.ptr(st.norms), GRAD_NORM, 1.0 / world_size,
                                 _lib.stream_ptr()))
    _lib.count_launches(3)


def apply_gradients(st, hp, world_size=1, allreduce=None, clipped=False):
    """model_helper.py:404-417: g += l2*w; per-tensor clip_by_norm(2); [mean over ranks]; Adam."""
    L = _lib.lib()
    if not clipped:
        regularise_and_clip(st, hp, world_size)
    if allreduce is not None:
        allreduce(st.grads)  # sum of (clipped / world_size) = CrossShardOptimizer's mean
    st._step += 1
    st.step_dev.add_(1)
    b1, b2, eps = 0.9, 0.999, 1e-8
    lr_t = float(hp["learning_rate"]) * (1.0 - b2 ** st.step) ** 0.5 / (1.0 - b1 ** st.step)
    _lib.check(L.plas_adam_step(_lib.ptr(st.params), _lib.ptr(st.grads), _lib.ptr(st.m), _lib.ptr(st.v), st.total, lr_t, b1, b2, eps,
                                1.0, _lib.stream_ptr()))
    _lib.count_launches(1)
    # --add_noise N (model_helper.py:418-432): when the global step (before this update) is a positive multiple of N, every
    # variable named '.../kernel' also receives N(0, noise_std) noise
    period, prev_step = int(hp.get("add_noise", 0) or 0), st.step - 1
    if period > 0 and prev_step > 0 and prev_step % period == 0:
        add_weight_noise(st, hp)


WEIGHT_NOISE_TID = 900  # + index of the variable


def add_weight_noise(st, hp):
    std = float(hp.get("noise_std", 0.1))
    for i, name in enumerate(st.names):
        if name.endswith("kernel"):
            n = int(np.prod(st.shapes[name]))
            _lib.check(_lib.lib().plas_add_normal_noise_f32(st.w(name), n, drop_seed(int(hp.get("dropout_seed", 0)), 0, WEIGHT_NOISE_TID + i),
                                                            _lib.ptr(st.step_dev), 0.0, std, _lib.stream_ptr()))
            _lib.count_launches(1)


def reference_weight_noise(st, hp, name, step):
    """numpy mirror of the noise ``add_weight_noise`` adds to variable ``name`` when the device step counter reads ``step``."""
    i, n = st.index[name], int(np.prod(st.shapes[name]))
    seed = drop_seed(int(hp.get("dropout_seed", 0)), step, WEIGHT_NOISE_TID + i)
    u1 = (hash_u24(n, seed).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    u2 = (hash_u24(n, (seed + 1) & 0xFFFFFFFF).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    z = np.sqrt(-2.0 * np.log(u1.astype(np.float64))) * np.cos((np.float32(6.283185307179586) * u2).astype(np.float64))
    return (float(hp.get("noise_std", 0.1)) * z).reshape(st.shapes[name])


def train_step(features, labels, st, hp, binf=None, world_size=1, allreduce=None):
    """One optimiser step.  Returns {'loss': audio_loss + L2 term, parts...} as device scalars."""
    parts = forward_backward(features, labels, st, hp, binf)
    apply_gradients(st, hp, world_size, allreduce)
    reg = st.wsq.sum() * (0.5 * float(hp.get("l2_reg_scale", 0.0)))  # L2 term of the PRE-update weights
    parts["loss"] = parts["audio_loss"] + reg
    return parts


class GraphedTrainStep:
    """The training step captured once in a CUDA graph (the ~500 small launches of a step -- per-step decoder kernels on
    two streams, GEMMs, persistent recurrences -- replay without host involvement).  Shapes (B, T, S) are those of the
    example batch; lengths and values may change freely between replays.  Adam's step-dependent learning rate and the
    data-parallel all-reduce stay outside the graph."""

    def __init__(self, features, labels, st, hp, binf=None, world_size=1, allreduce=None):
        self.st, self.hp, self.world, self.allreduce = st, hp, world_size, allreduce
        self.f = {k: v.clone() for k, v in features.items()}
        self.l = {k: v.clone() for k, v in labels.items()}
        self.binf = binf
        warm = torch.cuda.Stream(device=st.params.device)
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):  # one eager pass: function attributes set, allocator primed, side streams created
            forward_backward(self.f, self.l, st, hp, binf)
            regularise_and_clip(st, hp, world_size)
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count
        with torch.cuda.graph(self.graph):
            self.parts = forward_backward(self.f, self.l, st, hp, binf)
            regularise_and_clip(st, hp, world_size)
        self.launches_per_replay = _lib.launch_count - n0

    def __call__(self, features, labels):
        for k, v in features.items():
            self.f[k].copy_(v, non_blocking=True)
        for k, v in labels.items():
            self.l[k].copy_(v, non_blocking=True)
        self.graph.replay()
        _lib.count_launches(self.launches_per_replay)
        apply_gradients(self.st, self.hp, self.world, self.allreduce, clipped=True)
        parts = dict(self.parts)
        parts["loss"] = parts["audio_loss"] + self.st.wsq.sum() * (0.5 * float(self.hp.get("l2_reg_scale", 0.0)))
        return parts


def train_variable_shapes(hp, num_channels=None, binf_count=0):
    """variable_shapes + the 'speller_binf/' twin of the multitask configuration (model_helper.py:221)."""
    shapes = dict(variable_shapes(hp, num_channels))
    if hp.get("binary_outputs") and hp.get("binf_trainable") and hp.get("binf_projection"):
        shapes["binf2phone"] = (binf_count, hp["target_vocab_size"])  # model_helper.py:181-184: a variable initialised U(0, 1) (weights.init_params), not from the constant map
    if hp.get("binary_outputs"):
        V = hp["target_vocab_size"]
        for k, s in list(shapes.items()):
            if not k.startswith("speller/"):
                continue
            nk = "speller_binf/" + k[len("speller/"):]
            if k.endswith(("cell_0/lstm_cell/kernel", "cell_0_attention/attention_wrapper/lstm_cell/kernel")):
                s = (s[0] - V + binf_count, s[1])
            elif hp.get("binf_projection"):
                pass  # Dense(V) is built on the 2n-wide attention but never used (inner_projection_layer=False, las/model.py:251-257)
            elif k.endswith("projection_layer/kernel"):
                s = (s[0], binf_count)
            elif k.endswith("projection_layer/bias"):
                s = (binf_count,)
            shapes[nk] = s
        if not hp.get("multitask"):
            shapes = {k: s for k, s in shapes.items() if not k.startswith("speller/")}
    return shapes
