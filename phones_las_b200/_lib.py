"""ctypes binding of libplas.so (include/plas.h).  No fallback: a missing library raises."""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libplas.so")

PLAS_F32, PLAS_BF16 = 0, 1
ATT_CODES = {"luong": 0, "bahdanau": 1, "luong_monotonic": 2, "bahdanau_monotonic": 3, "custom": 4}


class PlasError(RuntimeError):
    pass


class FrontendDesc(C.Structure):
    _fields_ = [("backend", C.c_int32), ("feature_type", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32),
                ("n_mels", C.c_int32), ("n_mfcc", C.c_int32), ("energy", C.c_int32), ("deltas", C.c_int32),
                ("sp_delta_literal", C.c_int32), ("n_fac", C.c_int32), ("fac", C.c_int32 * 8),
                ("fb_total", C.c_int32), ("_pad", C.c_int32),
                ("window", C.c_void_p), ("tw", C.c_void_p), ("tw_unpack", C.c_void_p),
                ("fb_start", C.c_void_p), ("fb_len", C.c_void_p), ("fb_off", C.c_void_p), ("fb_w", C.c_void_p),
                ("dct", C.c_void_p), ("mean", C.c_void_p), ("stdv", C.c_void_p)]


class RecDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("B", C.c_int32), ("T", C.c_int32), ("U", C.c_int32), ("ndir", C.c_int32),
                ("out_zeroed", C.c_int32),
                ("xproj", C.c_void_p), ("whh", C.c_void_p), ("lengths", C.c_void_p), ("out", C.c_void_p),
                ("out_batch_stride", C.c_int64), ("c_final", C.c_void_p), ("h_final", C.c_void_p),
                ("whh_tc", C.c_void_p), ("max_clusters", C.c_int32), ("reserved0", C.c_int32)]


class DecDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("B", C.c_int32), ("Tm", C.c_int32), ("D", C.c_int32), ("Ud", C.c_int32),
                ("V", C.c_int32), ("n_layers", C.c_int32), ("attention_type", C.c_int32),
                ("sos_id", C.c_int32), ("eos_id", C.c_int32), ("max_steps", C.c_int32),
                ("teacher_forced", C.c_int32), ("decoding_length_factor", C.c_float), ("score_bias", C.c_float),
                ("keys", C.c_void_p), ("values", C.c_void_p), ("mem_len", C.c_void_p),
                ("w_cell", C.c_void_p * 4), ("w_emb", C.c_void_p), ("b_cell", C.c_void_p * 4),
                ("w_query", C.c_void_p), ("v_att", C.c_void_p), ("w_proj", C.c_void_p), ("b_proj", C.c_void_p),
                ("forced_ids", C.c_void_p), ("logits", C.c_void_p), ("sample_ids", C.c_void_p),
                ("alignment", C.c_void_p), ("seq_len", C.c_void_p), ("n_steps", C.c_void_p),
                ("w_cell_tc", C.c_void_p * 4), ("w_query_tc", C.c_void_p), ("pv", C.c_void_p),
                ("pv_ld", C.c_int32), ("_pad2", C.c_int32),
                ("vw", C.c_void_p), ("w_x_tc", C.c_void_p * 4), ("w_h_tc", C.c_void_p * 4)]


class GemmExDesc(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int32), ("A", C.c_void_p), ("sam", C.c_int64),
                ("sak", C.c_int64), ("B", C.c_void_p), ("sbk", C.c_int64), ("sbn", C.c_int64), ("C", C.c_void_p),
                ("ldc", C.c_int64), ("bias", C.c_void_p), ("alpha", C.c_float), ("beta", C.c_float),
                ("batch", C.c_int32), ("_pad", C.c_int32), ("batch_a", C.c_int64), ("batch_b", C.c_int64),
                ("batch_c", C.c_int64), ("split_ws", C.c_void_p), ("split_ws_bytes", C.c_size_t)]


class RecTrainDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("U", C.c_int32), ("ndir", C.c_int32), ("din", C.c_int32),
                ("_pad", C.c_int32), ("z", C.c_void_p), ("kernel", C.c_void_p * 2), ("lengths", C.c_void_p),
                ("out", C.c_void_p), ("out_batch_stride", C.c_int64), ("c_save", C.c_void_p), ("h_prev", C.c_void_p),
                ("dout", C.c_void_p), ("c_final", C.c_void_p), ("h_final", C.c_void_p), ("dc_final", C.c_void_p), ("dh_final", C.c_void_p)]


class DecTrainDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("S", C.c_int32), ("Tm", C.c_int32), ("D", C.c_int32), ("Ud", C.c_int32),
                ("E", C.c_int32), ("n_out", C.c_int32), ("n_layers", C.c_int32), ("attention_type", C.c_int32),
                ("dmemory_accumulate", C.c_int32), ("keep_prob", C.c_float), ("drop_seed", C.c_uint32),
                ("bottom_only", C.c_int32), ("_pad", C.c_int32),
                ("kernel", C.c_void_p * 4), ("bias", C.c_void_p * 4), ("w_mem", C.c_void_p), ("w_query", C.c_void_p),
                ("v_att", C.c_void_p), ("w_proj", C.c_void_p), ("b_proj", C.c_void_p), ("memory", C.c_void_p),
                ("mem_len", C.c_void_p), ("x_in", C.c_void_p), ("logits", C.c_void_p), ("dlogits", C.c_void_p),
                ("dkernel", C.c_void_p * 4), ("dbias", C.c_void_p * 4), ("dw_mem", C.c_void_p), ("dw_query", C.c_void_p),
                ("dv_att", C.c_void_p), ("dw_proj", C.c_void_p), ("db_proj", C.c_void_p), ("dmemory", C.c_void_p),
                ("drop_step", C.c_void_p), ("c_init", C.c_void_p * 4), ("h_init", C.c_void_p * 4),
                ("dc_init", C.c_void_p * 4), ("dh_init", C.c_void_p * 4),
                ("sample_prob", C.c_float), ("sample_seed", C.c_uint32), ("xdrop_seed", C.c_uint32), ("_pad2", C.c_uint32),
                ("x_in_rw", C.c_void_p), ("w_att_layer", C.c_void_p), ("dw_att_layer", C.c_void_p), ("att_layer", C.c_int32),
                ("_pad3", C.c_int32), ("score_bias", C.c_void_p), ("dscore_bias", C.c_void_p),
                ("sigmoid_noise", C.c_float), ("noise_seed", C.c_uint32), ("att_out", C.c_void_p), ("datt_extra", C.c_void_p),
                ("dx_in", C.c_void_p), ("sample_table", C.c_void_p), ("sample_fed_ids", C.c_void_p)]


class DecInferDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("Tm", C.c_int32), ("D", C.c_int32), ("Ud", C.c_int32), ("V", C.c_int32),
                ("n_layers", C.c_int32), ("attention_type", C.c_int32), ("sos_id", C.c_int32), ("eos_id", C.c_int32),
                ("max_steps", C.c_int32), ("teacher_forced", C.c_int32), ("decoding_length_factor", C.c_float),
                ("kernel", C.c_void_p * 4), ("bias", C.c_void_p * 4), ("w_query", C.c_void_p), ("v_att", C.c_void_p),
                ("w_proj", C.c_void_p), ("b_proj", C.c_void_p), ("keys", C.c_void_p), ("values", C.c_void_p),
                ("mem_len", C.c_void_p), ("forced_ids", C.c_void_p), ("logits", C.c_void_p), ("sample_ids", C.c_void_p),
                ("alignment", C.c_void_p), ("seq_len", C.c_void_p), ("n_steps", C.c_void_p),
                ("bottom_only", C.c_int32), ("att_layer", C.c_int32), ("c_init", C.c_void_p * 4), ("h_init", C.c_void_p * 4),
                ("w_att_layer", C.c_void_p), ("score_bias", C.c_void_p), ("beam_width", C.c_int32), ("_pad_beam", C.c_int32),
                ("beam_predicted", C.c_void_p), ("beam_parent", C.c_void_p), ("beam_word", C.c_void_p), ("beam_scores", C.c_void_p),
                ("beam_lengths", C.c_void_p)]


EXPORTS = {
    "plas_last_error": (C.c_char_p, []),
    "plas_version": (C.c_int, []),
    "plas_num_sms": (C.c_int, []),
    "plas_frontend_workspace_bytes": (C.c_size_t, [C.POINTER(FrontendDesc), C.c_int32, C.c_int32]),
    "plas_frontend_fwd": (C.c_int, [C.POINTER(FrontendDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "plas_gemm_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "plas_gemm_bf16_f32out": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32,
                                        C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "plas_gemm_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "plas_cast_pad_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_void_p]),
    "plas_rec_units_per_cta": (C.c_int32, [C.c_int32, C.c_int32]),
    "plas_rec_workspace_bytes": (C.c_size_t, [C.POINTER(RecDesc)]),
    "plas_bilstm_rec_fwd": (C.c_int, [C.POINTER(RecDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_relu_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "plas_add_normal_noise_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "plas_log_probs_reg_grad": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "plas_decoder_workspace_bytes": (C.c_size_t, [C.POINTER(DecDesc)]),
    "plas_decoder_fwd": (C.c_int, [C.POINTER(DecDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_seq_ce_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "plas_sigmoid_ce_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "plas_ctc_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "plas_mask_time": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_void_p]),
    "plas_gemm_f32_ex": (C.c_int, [C.POINTER(GemmExDesc), C.c_void_p]),
    "plas_colsum_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "plas_rec_train_workspace_bytes": (C.c_size_t, [C.POINTER(RecTrainDesc)]),
    "plas_bilstm_rec_train_fwd": (C.c_int, [C.POINTER(RecTrainDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_bilstm_rec_train_bwd": (C.c_int, [C.POINTER(RecTrainDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_dec_train_workspace_bytes": (C.c_size_t, [C.POINTER(DecTrainDesc)]),
    "plas_decoder_train_fwd": (C.c_int, [C.POINTER(DecTrainDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_decoder_train_bwd": (C.c_int, [C.POINTER(DecTrainDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_decoder_infer_f32_workspace_bytes": (C.c_size_t, [C.POINTER(DecInferDesc)]),
    "plas_decoder_infer_f32": (C.c_int, [C.POINTER(DecInferDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_seq_ce_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "plas_sigmoid_ce_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "plas_ctc_grad_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "plas_ctc_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                C.c_void_p]),
    "plas_grad_l2_norm_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "plas_grad_l2_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_clip_scale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "plas_dropout_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_float, C.c_void_p]),
    "plas_axpy_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "plas_split3_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_void_p]),
    "plas_gemm_tf32x3_scratch_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "plas_gemm_tf32x3_tn": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "plas_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_float, C.c_void_p]),
}

_lib = None


def lib():
    """Load libplas.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PlasError(f"{LIB_PATH} not found: run `python -m phones_las_b200.build` "
                            "(there is no CPU fallback)")
        from . import build as _build
        stamp = os.path.join(HERE, "csrc", ".libplas.stamp")
        if os.path.exists(stamp) and open(stamp).read() != _build._digest():
            raise PlasError(f"{LIB_PATH} is older than its sources: run `python -m phones_las_b200.build`")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise PlasError(f"libplas error {rc}: {lib().plas_last_error().decode()}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return C.c_void_p(t.data_ptr())


def require_cuda():
    if not torch.cuda.is_available():
        raise PlasError("phones_las_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib()


def dtype_code(precision):
    return {"fp32": PLAS_F32, "bf16": PLAS_BF16}[precision]


def torch_dtype(precision):
    return {"fp32": torch.float32, "bf16": torch.bfloat16}[precision]


# Kernels whose CTAs spin on each other beyond one thread-block cluster (grid barriers through L2: the fused decoders, the
# cooperative recurrences) need ALL their CTAs resident at once.  Two of them in flight on different streams could each hold
# part of the GPU and wait for the rest forever, so every such launch is chained behind the previous one with an event: at most
# one is ever in flight.  (rec_tc_kernel only synchronises inside its clusters; GEMMs / front-end never wait on other CTAs.)
_grid_sync_done = None


class grid_sync_kernel:
    """``with _lib.grid_sync_kernel():`` around a launch that needs device-wide co-residency."""

    def __enter__(self):
        if _grid_sync_done is not None:
            torch.cuda.current_stream().wait_event(_grid_sync_done)
        return self

    def __exit__(self, *exc):
        global _grid_sync_done
        ev = torch.cuda.Event()
        ev.record()
        _grid_sync_done = ev


# SM budget of the tcgen05 recurrence (-> plas_rec_desc.max_clusters; 0 = every cluster the GPU holds, the lowest latency): the
# serving loop lowers it while consecutive batches overlap on several streams, so that the other batch's kernels find free SMs
# (c2, two streams: 64 SMs = 4 clusters of 16 -> 86 k audio-s/s end to end, all 6 placeable clusters -> 81 k)
rec_sm_budget = 0


class rec_sms:
    """``with _lib.rec_sms(64):`` -- recurrences launched inside occupy at most that many SMs."""

    def __init__(self, n):
        self.n = int(n)

    def __enter__(self):
        global rec_sm_budget
        self.prev, rec_sm_budget = rec_sm_budget, self.n
        return self

    def __exit__(self, *exc):
        global rec_sm_budget
        rec_sm_budget = self.prev


# launch accounting for bench.py's gpu_launches (our kernels only)
launch_count = 0


def count_launches(n):
    global launch_count
    launch_count += n


# optional per-stage device timeline (bench.py's roofline leg): a list of (stage, start, stop)
# CUDA events recorded on torch's current stream -- the stream every kernel is launched on.
_timeline = None


def timeline_start():
    global _timeline
    _timeline = []


def timeline_stop():
    """-> {stage: [ms per launch, ...]} (synchronises)."""
    global _timeline
    tl, _timeline = _timeline, None
    torch.cuda.synchronize()
    out = {}
    for name, a, b in tl or []:
        out.setdefault(name, []).append(a.elapsed_time(b))
    return out


class stage:
    """``with _lib.stage("rec"):`` brackets one C-ABI call with CUDA events when a timeline is active."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _timeline is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _timeline is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _timeline.append((self.name, self.a, b))
        return False
