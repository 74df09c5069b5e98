/* plas.h -- C ABI of libplas.so: the B200-native (sm_100a) phones-las hot path.
 *
 * The reference (sciforce/phones-las) is pure Python/TensorFlow and has no FFI of its own;
 * each entry point below names the reference call it replaces (file:line in the reference
 * tree).  Conventions:
 *   - every function returns 0 on success or a negative PLAS_E* code; the message is
 *     available from plas_last_error() (thread-local);
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every
 *     buffer including workspaces (sizes from the *_workspace_bytes queries); nothing is
 *     allocated or cached behind the caller's back;
 *   - `stream` is a cudaStream_t (passed as void*); all work is enqueued on it, no call
 *     synchronises the device;
 *   - tensors are row-major, batch-major ([B,T,C]) like the reference's TF tensors;
 *   - dtype codes: PLAS_F32 = 0, PLAS_BF16 = 1.
 */
#ifndef PLAS_H_
#define PLAS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLAS_F32 0
#define PLAS_BF16 1

#define PLAS_ATT_LUONG 0
#define PLAS_ATT_BAHDANAU 1
#define PLAS_ATT_LUONG_MONOTONIC 2
#define PLAS_ATT_BAHDANAU_MONOTONIC 3 /* fp32 step kernels only (plas_decoder_infer_f32: mode 'hard'; plas_decoder_train_*: sigmoid noise) */
#define PLAS_ATT_CUSTOM 4             /* CustomAttention, las/model.py:72-101: fp32 step kernels, and plas_decoder_fwd's folded bf16 tensor-core kernel (keys already relu'd by the caller) */

typedef void* plas_stream_t;

const char* plas_last_error(void);
int plas_version(void);
int plas_num_sms(void);

/* ------------------------------------------------------------------------------------
 * K1  acoustic front-end.  Replaces calculate_acoustic_features (preprocess_all.py:69-130)
 * -> speechpy.feature.{mfe,mfcc,extract_derivative_feature} / librosa.feature.{melspectrogram,
 * mfcc,rms,delta} + amplitude_to_db, and the per-channel normalisation of
 * utils/dataset_utils.py:213-220, batched over utterances.
 * ---------------------------------------------------------------------------------- */
typedef struct plas_frontend_desc {
  int32_t backend;          /* 0 = speechpy, 1 = librosa (preprocess_all.py:204-205)            */
  int32_t feature_type;     /* 0 = mfe, 1 = mfcc        (preprocess_all.py:202-203)            */
  int32_t n_fft;            /* int(window * 16)         (preprocess_all.py:70)                 */
  int32_t hop;              /* int(step * 16)           (preprocess_all.py:71)                 */
  int32_t n_mels;
  int32_t n_mfcc;
  int32_t energy;           /* --energy                                                        */
  int32_t deltas;           /* --deltas                                                        */
  int32_t sp_delta_literal; /* speechpy derivative_extraction reading              */
  int32_t n_fac;            /* radix factorisation of n_fft/2 (radices in {2,3,4,5,8})         */
  int32_t fac[8];
  int32_t fb_total;         /* number of non-zero filterbank weights                           */
  int32_t _pad;
  const float* window;      /* [n_fft] analysis window (ones / periodic Hann)                  */
  const float* tw;          /* [n_fft/2][2]   exp(-2*pi*i*k/(n_fft/2))                         */
  const float* tw_unpack;   /* [n_fft/2+1][2] exp(-2*pi*i*k/n_fft)                             */
  const int32_t* fb_start;  /* [n_mels] first FFT bin of each mel filter                       */
  const int32_t* fb_len;    /* [n_mels] number of bins                                         */
  const int32_t* fb_off;    /* [n_mels] offset into fb_w                                       */
  const float* fb_w;        /* [fb_total] filter weights; rows whose fb_len and fb_off are
                               multiples of 4 (zero-padded by the host) take a float4 path     */
  const float* dct;         /* [n_mfcc][n_mels] orthonormal DCT-II rows (mfcc) or NULL         */
  const float* mean;        /* [C] or NULL  (utils/dataset_utils.py:213-220)                   */
  const float* stdv;        /* [C] or NULL                                                     */
} plas_frontend_desc;

size_t plas_frontend_workspace_bytes(const plas_frontend_desc* d, int32_t B, int32_t T_max);

/* wave [B][wave_stride] f32, n_samples [B]; feats [B][T_max][C] f32 (zero past each utterance's
 * frame count), n_frames [B] written by the kernel. */
int plas_frontend_fwd(const plas_frontend_desc* d, const float* wave, const int32_t* n_samples,
                      int32_t B, int64_t wave_stride, float* feats, int32_t* n_frames,
                      int32_t T_max, int32_t C, void* workspace, size_t workspace_bytes,
                      plas_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K2  time-parallel projections.  Replace the x-part of `concat([x,h]) @ kernel + bias`
 * inside tf.nn.rnn_cell.LSTMCell (las/ops.py:11-12,35-40) hoisted out of the time loop, the
 * pyramidal frame concat (las/ops.py:49-65, a free view here) and the attention
 * memory_layer Dense (las/model.py:168-169).
 *   C[M][ldc] = A[M][K] * Wt[N][K]^T + bias[N]
 * bf16: tcgen05/TMEM/TMA kernel (N % 128 == 0, lda/ldw multiples of 8, 16-byte aligned bases).
 * f32 : exact-fp32 SIMT kernel (reference-precision mode).
 * ---------------------------------------------------------------------------------- */
int plas_gemm_bf16(const void* A, int64_t M, int32_t K, int64_t lda, const void* Wt, int32_t N,
                   int64_t ldw, const float* bias, void* C, int64_t ldc, plas_stream_t stream);
/* same product with an f32 result (used for PV = values x projection kernel, las/model.py:251-257) */
int plas_gemm_bf16_f32out(const void* A, int64_t M, int32_t K, int64_t lda, const void* Wt, int32_t N,
                          int64_t ldw, const float* bias, float* C, int64_t ldc, plas_stream_t stream);
int plas_gemm_f32(const float* A, int64_t M, int32_t K, int64_t lda, const float* Wt, int32_t N,
                  int64_t ldw, const float* bias, float* C, int64_t ldc, plas_stream_t stream);
/* f32 [rows][ld_in] -> bf16 [rows][ld_out], columns >= cols zero-filled. */
int plas_cast_pad_bf16(const float* x, int64_t rows, int32_t cols, int64_t ld_in, void* y,
                       int64_t ld_out, plas_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K3  persistent (bi)directional LSTM recurrence.  Replaces tf.nn.bidirectional_dynamic_rnn /
 * tf.nn.dynamic_rnn over LSTMCell (las/ops.py:23-46) given the hoisted projections.
 * Gate order inside a unit is TF's (i, j, f, o); forget_bias 1.0 is applied at run time.
 * ---------------------------------------------------------------------------------- */
typedef struct plas_rec_desc {
  int32_t dtype;            /* PLAS_F32 / PLAS_BF16: type of xproj, whh, out                   */
  int32_t B, T, U, ndir;    /* T = time extent of xproj/out; ndir 1 (fw) or 2 (fw,bw)          */
  int32_t out_zeroed;       /* 1: the caller already zeroed `out`; 0: the call must leave out[b][t >= len] = 0 */
  const void* xproj;        /* [B][T][ndir*4U], column = dir*4U + 4*unit + gate                */
  const void* whh;          /* packed recurrent weights, see plas_rec_pack_whh                  */
  const int32_t* lengths;   /* [B]                                                             */
  void* out;                /* [B][T_out][ndir*U], zero for t >= len (see out_zeroed)          */
  int64_t out_batch_stride; /* elements between utterances in `out` (>= T*ndir*U)              */
  float* c_final;           /* [ndir][B][U]                                                    */
  float* h_final;           /* [ndir][B][U]                                                    */
  const void* whh_tc;       /* bf16, optional: [ndir][U/32][128][U], row (TMEM lane) 32*q + 8*gate + u8 holds
                               W_h[:, gate*U + 32*cta + 8*q + u8]; the TMEM-resident operand of the tcgen05
                               recurrence (NULL: mma.sync kernels)                                         */
  int32_t max_clusters;     /* tcgen05 recurrence: upper bound on the thread-block clusters (of U/32 CTAs) the call may
                               occupy; 0 = as many as the GPU holds at once (lowest latency).  A serving loop that overlaps
                               consecutive batches on several streams passes a smaller budget so that the other batch's
                               kernels find free SMs                                                            */
  int32_t reserved0;
} plas_rec_desc;

/* Layout contract for whh (host side packs once per checkpoint load):
 *   f32 : [ndir][U/upc][U(k)][4*upc]  with column = 4*unit_local + gate
 *   bf16: [ndir][U/32][8 warps][U/16 ksteps][32 lanes][8 bf16] mma.m16n8k16 B-fragments
 * plas_rec_units_per_cta reports upc for (dtype,U).
 * Three kernels sit behind plas_bilstm_rec_fwd: tcgen05 with W_hh resident in tensor memory (bf16,
 * U in {64,128,256,512}, needs whh_tc), an mma.sync thread-block-cluster kernel with W_hh resident in
 * registers (bf16, same widths) and an L2-exchange cooperative kernel (any width, f32 or bf16). */
int32_t plas_rec_units_per_cta(int32_t dtype, int32_t U);
size_t plas_rec_workspace_bytes(const plas_rec_desc* d);
int plas_bilstm_rec_fwd(const plas_rec_desc* d, void* workspace, size_t workspace_bytes,
                        plas_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K4  attention decoder.  Replaces AttentionWrapper(MultiRNNCell) + BasicDecoder +
 * GreedyEmbeddingHelper / TrainingHelper + dynamic_decode (las/model.py:145-202, 205-349) and the
 * DenseBinfDecoder projection (utils/training_helper.py:122-153), all steps on the device.
 * Two kernels sit behind plas_decoder_fwd: a SIMT/mma.sync kernel (any shape, f32 or bf16) and a
 * weight-stationary TMA + tcgen05 kernel (bf16, B <= 128, D and Ud multiples of 64) chosen when the
 * *_tc weight layouts and pv are supplied.
 * ---------------------------------------------------------------------------------- */
typedef struct plas_dec_desc {
  int32_t dtype;            /* PLAS_F32 / PLAS_BF16: type of keys, values, packed weights      */
  int32_t B, Tm, D, Ud, V, n_layers, attention_type;
  int32_t sos_id, eos_id;
  int32_t max_steps;        /* capacity of the output buffers along the step axis              */
  int32_t teacher_forced;   /* 0 = greedy, 1 = feed forced_ids, run exactly max_steps          */
  float decoding_length_factor; /* greedy: stop at rint(max(mem_len) * factor)                 */
  float score_bias;         /* luong_monotonic attention_score_bias                            */
  const void* keys;         /* [B][Tm][Ud]  memory_layer(values)                               */
  const void* values;       /* [B][Tm][D]   length-masked encoder outputs                      */
  const int32_t* mem_len;   /* [B]                                                             */
  const void* w_cell[4];    /* packed LSTM kernels (non-embedding rows), plas_dec_* layout     */
  const void* w_emb;        /* [V][4Ud] cell-0 kernel rows of the one-hot input, col = 4u+g    */
  const float* b_cell[4];   /* [4Ud] biases, col = 4*unit + gate                               */
  const void* w_query;      /* [Ud][Ud] bahdanau query_layer kernel (row-major k,u) or NULL    */
  const float* v_att;       /* [Ud] bahdanau attention_v or NULL                               */
  const void* w_proj;       /* [D][V] projection kernel                                        */
  const float* b_proj;      /* [V]                                                             */
  const int32_t* forced_ids;/* [B][max_steps] teacher-forced inputs or NULL                    */
  float* logits;            /* [B][max_steps][V]                                               */
  int32_t* sample_ids;      /* [B][max_steps]                                                  */
  float* alignment;         /* [B][max_steps][Tm] or NULL                                      */
  int32_t* seq_len;         /* [B] final_sequence_length                                       */
  int32_t* n_steps;         /* [1] number of decode iterations executed                        */
  /* tensor-core path (bf16; optional -- NULL selects the SIMT kernel).  Layouts: plas_dec_pack notes */
  const void* w_cell_tc[4]; /* [Ud/4][K_l/64][16 x 64 bf16, 128B-swizzled] UMMA B tiles per layer      */
  const void* w_query_tc;   /* [Ud/16][Ud/64][16 x 64 bf16, 128B-swizzled] bahdanau query_layer        */
  const float* pv;          /* [B][Tm][pv_ld] values x projection kernel (f32), logits = a.pv + b_proj */
  int32_t pv_ld;
  int32_t _pad2;
  /* folded-context tensor-core path (decoder_fold.cu; bf16, optional).  The context fed back to cell 0
   * (AttentionWrapper, las/model.py:195-200) is linear in the values, so vw = values x W0[V:V+D] lets the
   * attention phase emit cell 0's pre-activation part sum_t a_t vw[t] directly; the cells then only
   * contract over h (K = Ud).  Tile layout of w_x_tc / w_h_tc = w_cell_tc's with K = Ud.              */
  const void* vw;           /* [B][Tm][4Ud] bf16, column = 4*unit + gate                                */
  const void* w_x_tc[4];    /* l >= 1: rows 0..Ud-1 of cell l's kernel (its input = h of the cell below) */
  const void* w_h_tc[4];    /* recurrent rows of cell l's kernel (l = 0: rows V+D.. ; l >= 1: rows Ud..) */
} plas_dec_desc;

size_t plas_decoder_workspace_bytes(const plas_dec_desc* d);
int plas_decoder_fwd(const plas_dec_desc* d, void* workspace, size_t workspace_bytes,
                     plas_stream_t stream);

/* ------------------------------------------------------------------------------------
 * K5  loss forward passes of las_model_fn (model_helper.py:20-130, 347-358).  f32 in, f32 out; sums are
 * reduced in a fixed order (deterministic).  out3 = {sum(ce*w)/(sum(w)+1e-12), sum(ce*w), sum(w)}.
 * ---------------------------------------------------------------------------------- */
/* tf.contrib.seq2seq.sequence_loss (model_helper.py:30,75): logits [n_tokens][V], targets [n_tokens],
 * weights [n_tokens] (NULL = all ones); ce_tokens [n_tokens] receives the per-token cross-entropy. */
int plas_seq_ce_fwd(const float* logits, const int32_t* targets, const float* weights, int64_t n_tokens,
                    int32_t V, float* ce_tokens, float* out3, plas_stream_t stream);
/* sequence_loss_sigmoid (model_helper.py:81-95): per token the mean over n_feat of
 * tf.nn.sigmoid_cross_entropy_with_logits, then the weighted mean over tokens. */
int plas_sigmoid_ce_fwd(const float* logits, const float* labels, const float* weights, int64_t n_tokens,
                        int32_t n_feat, float* ce_tokens, float* out3, plas_stream_t stream);
/* tf.nn.ctc_loss_v2, dense labels, batch-major logits [B][T][C] (model_helper.py:355-356; blank = 0 there):
 * loss[b] = -log p(labels[b,:label_len[b]] | logits[b,:logit_len[b]]). */
int plas_ctc_fwd(const float* logits, const int32_t* labels, const int32_t* label_len, const int32_t* logit_len,
                 int32_t B, int32_t T, int32_t C, int32_t Lmax, int32_t blank, float* loss, plas_stream_t stream);

/* mask values past each length: y[b][t][:] = t < len[b] ? x[b][t][:] : 0  (attention `values`,
 * tf.contrib.seq2seq _prepare_memory; las/model.py:168-169). dtype-generic, in place allowed. */
int plas_mask_time(int32_t dtype, const void* x, void* y, const int32_t* len, int32_t B, int32_t T,
                   int32_t D, plas_stream_t stream);

/* ====================================================================================
 * TRAINING PATH (exact fp32, TF checkpoint layouts).  Replaces the TRAIN graph of las_model_fn
 * (model_helper.py:165-227, 319-358, 403-417): listener and speller forward with saved activations, their
 * gradients (optimizer.compute_gradients, model_helper.py:415), the loss heads' gradients, L2 regularisation,
 * per-tensor clip_by_norm and Adam.  Every weight and every weight gradient is in the layout of the TF variable
 * (LSTMCell kernel [din+U][4U] with gate column blocks i|j|f|o, Dense kernels [in][out]).
 * ==================================================================================== */

/* C[z] = alpha * A(m,k) B(k,n) + beta * C + bias[n] with arbitrary element strides:
 *   A(m,k) = A[z*batch_a + m*sam + k*sak],  B(k,n) = B[z*batch_b + k*sbk + n*sbn],  C[z*batch_c + m*ldc + n].
 * Serves x @ W, dZ @ W^T and X^T @ dZ on the TF layouts without transposed copies. */
typedef struct plas_gemm_ex_desc {
  int64_t M;
  int32_t N, K;
  const float* A;
  int64_t sam, sak;
  const float* B;
  int64_t sbk, sbn;
  float* C;
  int64_t ldc;
  const float* bias;        /* [N] or NULL */
  float alpha, beta;
  int32_t batch;            /* >= 1 independent problems (grid.z) */
  int32_t _pad;
  int64_t batch_a, batch_b, batch_c;
  float* split_ws;          /* optional scratch: long reductions over few output tiles (weight gradients) are cut
                               into K slices whose partial products are summed in a fixed order (deterministic) */
  size_t split_ws_bytes;
} plas_gemm_ex_desc;
int plas_gemm_f32_ex(const plas_gemm_ex_desc* d, plas_stream_t stream);
/* The same contractions on the tensor pipe with fp32-level accuracy (3xTF32, north_star item 2): every fp32 operand is split
 * into hi = nearest TF32 and lo = x - hi, and  a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo  is ONE tcgen05 kind::tf32 GEMM over a
 * concatenated contraction axis, [A_hi | A_lo | A_hi] . [B_hi | B_hi | B_lo]^T (gemm_tf32.cu).
 * plas_split3_f32 writes such an operand: out[r][seg*seg_ld + c] for seg = 0..2 from X[r][c] (transpose = 0, seg_ld >= cols), or
 * out[c][seg*seg_ld + r] (transpose = 1, seg_ld >= rows: turns X^T dZ / x W on the TF layouts into the TN form below); columns
 * past the data are zero; pattern 0 = (hi, lo, hi) for the left operand, 1 = (hi, hi, lo) for the right one.
 * plas_gemm_tf32x3_tn: C[M][N] (+)= A3[M][K3] . B3[N][K3]^T (+ bias[N]); K3 = 3*seg_ld; lda/ldb/ldc multiples of 4, 16-byte bases.
 * Problems with few output tiles and a long contraction (weight gradients) are cut into K slices whose partial tiles go to the
 * caller's scratch and are added in a fixed order (deterministic). */
int plas_split3_f32(const float* X, int64_t rows, int32_t cols, int64_t ld, float* out, int64_t ld_out, int64_t seg_ld,
                    int32_t pattern, int32_t transpose, plas_stream_t stream);
size_t plas_gemm_tf32x3_scratch_bytes(int64_t M, int32_t N, int32_t K3); /* split-K scratch (few output tiles, long K); 0 = none */
int plas_gemm_tf32x3_tn(const float* A3, int64_t M, int32_t K3, int64_t lda, const float* B3, int32_t N, int64_t ldb,
                        const float* bias, float* C, int64_t ldc, int32_t accumulate, void* scratch, size_t scratch_bytes,
                        plas_stream_t stream);
/* out[n] (+)= sum_m X[m][n]: bias gradients. */
int plas_colsum_f32(const float* X, int64_t M, int32_t N, int64_t ld, float* out, int32_t accumulate,
                    plas_stream_t stream);

/* (bi)directional LSTM recurrence of one listener layer, forward-with-save and BPTT (las/ops.py:23-46 and its
 * gradient).  `z` [B][T][ndir][4U] holds x W_x + b on entry of the forward call, the activated gates
 * (sigmoid i, tanh j, sigmoid(f+1), sigmoid o) on exit, and dL/dz (zero for t >= len) after the backward call. */
typedef struct plas_rec_train_desc {
  int32_t B, T, U, ndir, din;
  int32_t _pad;
  float* z;
  const float* kernel[2];   /* TF LSTMCell kernel [din+U][4U] of the fw / bw cell; rows din.. are W_hh          */
  const int32_t* lengths;   /* [B]                                                                            */
  float* out;               /* fwd: [B][T_out][ndir*U], caller-zeroed (stays 0 for t >= len)                  */
  int64_t out_batch_stride;
  float* c_save;            /* [B][T][ndir*U] cell states (fwd out, bwd in); NULL in a forward-only (inference) call */
  float* h_prev;            /* [B][T][ndir*U] h_{s-1} stored at the time index of step s (fwd out), caller-zeroed; may be NULL */
  const float* dout;        /* bwd: dL/dout, strides of `out`                                                 */
  float* c_final;           /* fwd, optional: [ndir][B][U] final cell / hidden states (encoder_state, las/ops.py:35-46) */
  float* h_final;
  const float* dc_final;    /* bwd, optional: gradients wrt the final states [ndir][B][U] (decoder cells seeded from them, */
  const float* dh_final;    /*                pass_hidden_state, las/model.py:259-267)                                      */
} plas_rec_train_desc;
size_t plas_rec_train_workspace_bytes(const plas_rec_train_desc* d);
int plas_bilstm_rec_train_fwd(const plas_rec_train_desc* d, void* workspace, size_t workspace_bytes, plas_stream_t stream);
int plas_bilstm_rec_train_bwd(const plas_rec_train_desc* d, void* workspace, size_t workspace_bytes, plas_stream_t stream);

/* Teacher-forced speller, forward and backward (las/model.py:205-296 TRAIN branch with TrainingHelper /
 * TrainingSigmoidHelper, sampling_probability = 0; attention luong or bahdanau, default wiring).  x_in is the
 * embedded decoder input of every step: one-hot ids (embedding_fn, las/model.py:245-246) or the binary-feature
 * vectors of the binary_outputs speller (las/model.py:237-241).  The workspace carries the saved activations from
 * the forward to the backward call. */
typedef struct plas_dec_train_desc {
  int32_t B, S, Tm, D, Ud, E, n_out, n_layers, attention_type;
  int32_t dmemory_accumulate; /* 1: dmemory += ..., 0: dmemory = ...                                           */
  float keep_prob;          /* 1 - dropout of the decoder cells' inputs (1.0 = off).  x_in must already be dropped out by
                               the caller (plas_dropout_f32); the kernels drop attention_{t-1} and the inter-layer h     */
  uint32_t drop_seed;       /* masks: attention uses drop_seed, layer l's output uses drop_seed + 1 + l                  */
  int32_t bottom_only;      /* AttentionMultiCell wiring (las/model.py:20-69,185-193): cell l >= 1 kernels are
                               [(l == 1 ? D : Ud) + D + Ud][4Ud], w_proj is [Ud][n_out] when n_layers > 1; keep_prob must be 1 */
  int32_t _pad;
  const float* kernel[4];   /* cell_k/lstm_cell/kernel [(k == 0 ? E + D : Ud) + Ud][4Ud]                       */
  const float* bias[4];     /* [4Ud]                                                                          */
  const float* w_mem;       /* memory_layer/kernel [D][Ud]                                                    */
  const float* w_query;     /* bahdanau query_layer/kernel [Ud][Ud] or NULL                                   */
  const float* v_att;       /* bahdanau attention_v [Ud] or NULL                                              */
  const float* w_proj;      /* projection_layer/kernel [D][n_out]                                             */
  const float* b_proj;      /* [n_out]                                                                        */
  const float* memory;      /* [B][Tm][D] encoder outputs, zero for t >= mem_len                              */
  const int32_t* mem_len;   /* [B]                                                                            */
  const float* x_in;        /* [B][S][E]                                                                      */
  float* logits;            /* fwd out: [B][S][n_out]                                                         */
  const float* dlogits;     /* bwd in : [B][S][n_out]                                                         */
  float* dkernel[4];        /* bwd out, layouts of the parameters                                             */
  float* dbias[4];
  float* dw_mem;
  float* dw_query;
  float* dv_att;
  float* dw_proj;
  float* db_proj;
  float* dmemory;           /* bwd out: [B][Tm][D] gradient wrt the encoder outputs                           */
  const uint32_t* drop_step; /* device optimiser-step counter added to the dropout seeds (NULL = 0)             */
  const float* c_init[4];   /* bottom_only + pass_hidden_state: initial (c, h) of cell l [B][Ud] or NULL (zeros)  */
  const float* h_init[4];
  float* dc_init[4];        /* bwd out, optional: gradients wrt the initial states                                */
  float* dh_init[4];
  /* scheduled sampling of the phone speller (las/model.py:279-288, utils/training_helper.py:48-87): with probability
   * sample_prob per (utterance, step) the next input is the one-hot of an id drawn from Categorical(logits_t); the forward call
   * rewrites x_in[b][t+1] through x_in_rw (= x_in) so that the backward call sees the inputs actually fed */
  float sample_prob;        /* 0 = teacher forcing                                                              */
  uint32_t sample_seed;     /* selection uses sample_seed, the categorical draw sample_seed + 1 (+ step, drop_step) */
  uint32_t xdrop_seed;      /* dropout seed of x_in (a sampled input is dropped out like the one it replaces)    */
  uint32_t _pad2;
  float* x_in_rw;
  /* attention_layer_size = att_layer (0 = none; default wiring, no dropout): attention = [h_top; context] W, W = w_att_layer
   * [Ud + D][A]; cell 0's kernel is then [E + A + Ud][4Ud] and w_proj [A][n_out] */
  const float* w_att_layer;
  float* dw_att_layer;
  int32_t att_layer;
  int32_t _pad3;
  /* luong_monotonic (tf.contrib.seq2seq.LuongMonotonicAttention, las/model.py:157-158): attention_score_bias [1] and its gradient */
  const float* score_bias;
  float* dscore_bias;
  /* bahdanau_monotonic in TRAIN mode (las/model.py:159-162): sigmoid_noise * N(0,1) is added to the scores; the normal deviates
   * come from the counter hash with seed noise_seed (+ drop_step), so the oracle can replay them (train.reference_noise) */
  float sigmoid_noise;
  uint32_t noise_seed;
  /* --binf_projection (las/model.py:180-183,251-257; utils/training_helper.py:17-27,122-153): the projection is the constant
   * [M; 1 - M] (w_proj; b_proj = zeros; dw_proj = db_proj = NULL) applied to the 2n-wide attention vectors; the regulariser
   * compute_log_probs_loss reads those vectors (att_out, fwd out [B][S][A], optional) and its gradient enters through
   * datt_extra (bwd in [B][S][A], optional).  Default wiring only. */
  float* att_out;
  const float* datt_extra;
  /* bwd out, optional: gradient wrt x_in [B][S][E] (embedding_size != 0, las/model.py:230-237: it flows into target_embedding) */
  float* dx_in;
  /* scheduled sampling when the decoder inputs are not one-hot (embedding_size != 0, --binf_projection): the sampled id feeds
   * row id of sample_table [n_out][E]; sample_fed_ids [B][S] (caller pre-fills it with targets_inputs) receives the ids fed */
  const float* sample_table;
  int32_t* sample_fed_ids;
} plas_dec_train_desc;
size_t plas_dec_train_workspace_bytes(const plas_dec_train_desc* d);
int plas_decoder_train_fwd(const plas_dec_train_desc* d, void* workspace, size_t workspace_bytes, plas_stream_t stream);
int plas_decoder_train_bwd(const plas_dec_train_desc* d, void* workspace, size_t workspace_bytes, plas_stream_t stream);

/* fp32 inference decoder (reference-precision mode) built from the same step kernels, on the TF checkpoint layout: greedy
 * (GreedyEmbeddingHelper + dynamic_decode, las/model.py:337-347) or teacher-forced, luong / bahdanau attention, default
 * wiring.  A host loop of small launches over a 2-slot state ring; launches after every utterance has finished are no-ops.
 * keys = memory_layer(values) is computed by the caller (plas_gemm_f32 / plas_gemm_f32_ex). */
typedef struct plas_dec_infer_desc {
  int32_t B, Tm, D, Ud, V, n_layers, attention_type, sos_id, eos_id;
  int32_t max_steps;        /* capacity of the outputs along the step axis                                     */
  int32_t teacher_forced;   /* 0 = greedy, 1 = feed forced_ids and run exactly max_steps                        */
  float decoding_length_factor; /* greedy: stop at rint(max(mem_len) * factor)                                 */
  const float* kernel[4];   /* cell_k/lstm_cell/kernel, TF layout [(k == 0 ? V + D : Ud) + Ud][4Ud]             */
  const float* bias[4];     /* [4Ud], gate blocks i|j|f|o                                                      */
  const float* w_query;     /* bahdanau query_layer/kernel [Ud][Ud] or NULL                                    */
  const float* v_att;       /* bahdanau attention_v [Ud] or NULL                                               */
  const float* w_proj;      /* projection_layer/kernel [D][V]                                                  */
  const float* b_proj;      /* [V]                                                                             */
  const float* keys;        /* [B][Tm][Ud]                                                                     */
  const float* values;      /* [B][Tm][D], zero for t >= mem_len                                               */
  const int32_t* mem_len;   /* [B]                                                                             */
  const int32_t* forced_ids;/* [B][max_steps] or NULL                                                          */
  float* logits;            /* [B][max_steps][V], caller-zeroed (steps that are not executed stay 0)           */
  int32_t* sample_ids;      /* [B][max_steps], caller-zeroed                                                   */
  float* alignment;         /* [B][max_steps][Tm] or NULL                                                      */
  int32_t* seq_len;         /* [B] final_sequence_length                                                       */
  int32_t* n_steps;         /* [1] number of decode iterations executed                                        */
  int32_t bottom_only;      /* AttentionMultiCell wiring (las/model.py:20-69,185-193): attention wraps cell 0 only; cell l >= 1 reads
                               [output below (the NEW attention for l = 1); OLD attention; h_{t-1}], kernel [(D or Ud) + D + Ud][4Ud];
                               the projection reads the top cell's h ([Ud][V]) when n_layers > 1                  */
  int32_t att_layer;        /* attention_layer_size A (0 = none): attention = [cell output; context] W, W = w_att_layer [Ud + D][A]; the
                               fed-back attention, cell 0's kernel ([V + A + Ud][4Ud]) and w_proj ([A][V]) then use A          */
  const float* c_init[4];   /* pass_hidden_state (las/model.py:259-267): initial cell / hidden state of layer l [B][Ud] or NULL */
  const float* h_init[4];
  const float* w_att_layer; /* attention_wrapper/attention_layer/kernel or NULL                                 */
  const float* score_bias;  /* luong_monotonic attention_score_bias [1] (device) or NULL                       */
  /* Beam search (PREDICT with beam_width > 0, las/model.py:219-226,298-319: tf.contrib.seq2seq.BeamSearchDecoder, length penalty
   * 0, + gather_tree).  B, keys, values, mem_len, c_init / h_init are then the TILED batch (tile_batch: row b*W + w); logits,
   * sample_ids and alignment are not written; seq_len [B] = dynamic_decode's per-hypothesis sequence lengths. */
  int32_t beam_width;       /* W (0 = greedy / teacher forced), <= 32                                          */
  int32_t _pad_beam;
  int32_t* beam_predicted;  /* out [B / W][max_steps][W]: predicted_ids after gather_tree (eos past each end)  */
  int32_t* beam_parent;     /* out [max_steps][B]: parent beam of every step                                   */
  int32_t* beam_word;       /* out [max_steps][B]: word id of every step                                       */
  float* beam_scores;       /* out [B] final log-probabilities, or NULL                                        */
  int32_t* beam_lengths;    /* out [B] final BeamSearchDecoderState.lengths, or NULL                           */
} plas_dec_infer_desc;
/* x = max(x, 0) in place: CustomAttention's keys = relu(memory_layer(values)) (las/model.py:94), applied by the caller of
 * plas_decoder_infer_f32 to the keys it passes. */
int plas_relu_f32(float* x, int64_t n, plas_stream_t stream);
size_t plas_decoder_infer_f32_workspace_bytes(const plas_dec_infer_desc* d);
int plas_decoder_infer_f32(const plas_dec_infer_desc* d, void* workspace, size_t workspace_bytes, plas_stream_t stream);

/* Loss heads with gradients.  dlogits = gscale * d(loss)/d(logits); out3 as in the forward-only calls. */
int plas_seq_ce_grad(const float* logits, const int32_t* targets, const float* weights, int64_t n_tokens, int32_t V,
                     float gscale, float* ce_tokens, float* out3, float* dlogits, plas_stream_t stream);
int plas_sigmoid_ce_grad(const float* logits, const float* labels, const float* weights, int64_t n_tokens,
                         int32_t n_feat, float gscale, float* ce_tokens, float* out3, float* dlogits,
                         plas_stream_t stream);
/* per-utterance CTC loss [B] and dlogits [B][T][C] = gscale * d(loss[b])/d(logits[b]) (zero for t >= logit_len) */
/* x[i] += mean + std * N(0,1): the periodic weight noise of model_helper.py:418-432 (NOISE_MEAN = 0, --noise_std) on one
 * 'kernel' variable; deviates from the counter hash with seed + *step * 0x85EBCA77 (step: device counter or NULL). */
int plas_add_normal_noise_f32(float* x, int64_t n, uint32_t seed, const uint32_t* step, float mean, float std, plas_stream_t stream);
/* compute_log_probs_loss (model_helper.py:132-146) and its gradient: att [n_rows][2 n_feat] = [log p1 | log p0];
 * out3[0] = weight * mean(|p1 + p0 - 1| + relu(log p1) + relu(log p0)); datt = d(out3[0]) / d(att); reg_rows [n_rows] scratch. */
int plas_log_probs_reg_grad(const float* att, int64_t n_rows, int32_t n_feat, float weight, float* reg_rows, float* out3,
                            float* datt, plas_stream_t stream);
size_t plas_ctc_grad_workspace_bytes(int32_t B, int32_t T, int32_t Lmax);
int plas_ctc_grad(const float* logits, const int32_t* labels, const int32_t* label_len, const int32_t* logit_len,
                  int32_t B, int32_t T, int32_t C, int32_t Lmax, int32_t blank, float gscale, float* loss,
                  float* dlogits, void* workspace, size_t workspace_bytes, plas_stream_t stream);

/* Optimiser (model_helper.py:404-417) on flat fp32 buffers; offsets [n_tensors+1] (device, int64) delimit the
 * variables.  grad_l2_norm: g += l2_scale*w, norms[i] = ||g_i||, wsq[i] = ||w_i||^2 (optional, for the reported L2 term);  clip_scale: g_i *= clip/max(norms[i],clip) * post_scale
 * (post_scale = 1/world_size before the data-parallel all-reduce);  adam_step: TF epsilon-hat form with
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller, gradients pre-multiplied by grad_scale. */
size_t plas_grad_l2_norm_scratch_bytes(int32_t n_tensors);
int plas_grad_l2_norm(const float* params, float* grads, const int64_t* offsets, int32_t n_tensors, float l2_scale,
                      float* norms, float* wsq, void* scratch, size_t scratch_bytes, plas_stream_t stream);
int plas_clip_scale(float* grads, const int64_t* offsets, int32_t n_tensors, const float* norms, float clip,
                    float post_scale, plas_stream_t stream);
/* Input dropout of the TRAIN graph (DropoutWrapper(input_keep_prob = 1 - dropout), las/ops.py:14-18): y = x * m / keep_prob
 * with a counter-based mask m(seed + *step * 0x85EBCA77, element index) -- `step` (device, may be NULL = 0) is the optimiser step,
 * read on the device so that a captured CUDA graph draws fresh masks on every replay; the same call on a gradient tensor is the
 * backward pass.  In place allowed. */
int plas_dropout_f32(const float* x, float* y, int64_t n, uint32_t seed, const uint32_t* step, float keep_prob,
                     plas_stream_t stream);
/* y += alpha * x: sums the encoder-output gradients of the heads (the spellers run on separate streams). */
int plas_axpy_f32(float* y, const float* x, int64_t n, float alpha, plas_stream_t stream);
int plas_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr_t, float beta1,
                   float beta2, float eps, float grad_scale, plas_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PLAS_H_ */
