/*
 * speechless_b200.h — C-ABI of libspeechless_b200.so
 *
 * Drop-in boundary for the wav2letter hot path of juliuskunze/speechless
 * (speechless/net.py::Wav2Letter).  The reference has no FFI layer of its own:
 * below the Python class surface it hands numpy batches to Keras/TensorFlow.
 * Each entry point below replaces one Keras/TF call site; the citation names
 * the reference line whose arithmetic the entry point reproduces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends
 *    in `_host`; the library allocates nothing on the hot path;
 *  - `stream` is a cudaStream_t passed as void*; calls are stream ordered and
 *    re-entrant;
 *  - return value 0 = ok, non-zero = error (message via sl_last_error);
 *    nothing throws or aborts across the ABI;
 *  - activations are channels-last (B, T, C) like Keras' Conv1D
 *    (net.py:304-305).  The tensor-core kernels read/write the *packed* form:
 *    16-bit elements (bf16, or fp16 in SL_PREC_FP16), channels zero-padded to a
 *    multiple of 64, optionally followed by a second "lo" plane (x - bf16(x)) in
 *    the same row for the split-bf16 (fp32-parity) mode: row = [hi(C_pad) | lo(C_pad)].
 */
#ifndef SPEECHLESS_B200_H
#define SPEECHLESS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SL_OK 0
#define SL_ERR_INVALID 1   /* bad shape / argument (host-detected)        */
#define SL_ERR_CUDA 2      /* CUDA runtime / driver error                 */
#define SL_ERR_INFEASIBLE 3 /* CTC label not alignable in the given frames */

/* precision of the packed 16-bit tensors (every tensor of one call chain uses the same one) */
#define SL_PREC_BF16 1    /* one bf16 plane, fp32 accumulate: 8 mantissa bits, fp32 range         */
#define SL_PREC_BF16X2 2  /* hi+lo bf16 planes, 3-term product (~fp32 parity), 3 MMAs per product  */
#define SL_PREC_FP16 3    /* one fp16 plane, fp32 accumulate: 11 mantissa bits at the cost of the
                             bf16 mode (logits <= 1e-3 rel at reference widths); gradients need a
                             power-of-two loss scale: pass it in sl_ctc_loss' grad_scale and its
                             inverse in sl_conv1d_wgrad's out_scale                                 */

/* epilogue / activation selector of sl_conv1d_fwd (net.py:298,304-305,328-330) */
#define SL_ACT_NONE 0
#define SL_ACT_RELU 1
#define SL_ACT_SOFTMAX 2 /* output_conv: bias + softmax, fp32 outputs */

int sl_version(void);
/* copies the calling thread's last error message (NUL terminated) into buf */
int sl_last_error(char* buf, size_t n);
/* device sync + sticky async error check (surfaces device-side traps) */
int sl_sync_check(void);

/* ---- packing between the Keras-facing fp32 tensors and the packed bf16 form ---- */

/* net.py:583-585 `_input_batch_and_prediction_lengths` feeds a zero padded
 * (B,T,F) float batch; this casts it to packed bf16 rows of `c_pad` channels
 * (T_alloc >= T rows per utterance, rows T..T_alloc zero).  */
int sl_pack_activation(const float* x, void* x_packed, int B, int T, int C,
                       int T_alloc, int c_pad, int prec, void* stream);
/* Raw-wave front layer (net.py:310-312 `wave_conv`: Conv1D k=250, strides=160,
 * padding="same", with the training-phase Dropout of net.py:301-303 in front):
 * lays the receptive field of every output frame out as one packed row,
 * y[b][t][j*C + c] = x[b][t*stride + j - pad_l][c] (TF SAME padding, zeros
 * outside), T_out = ceil(T/stride) rows of c_pad >= k*C channels, so that the
 * layer runs as a 1-tap sl_conv1d_fwd / sl_conv1d_wgrad over k*C channels.
 * drop_p > 0 applies inverted dropout per source sample (hash of seed). */
int sl_window_activation(const float* x, void* x_windowed, int B, int T, int C,
                         int k, int stride, int c_pad, int prec, float drop_p,
                         unsigned long long seed, void* stream);
/* inverse (debug / tests): packed -> fp32 (B,T,C) (hi + lo) */
int sl_unpack_activation(const void* x_packed, float* x, int B, int T, int C,
                         int T_alloc, int c_pad, int prec, void* stream);

/* Keras kernel layout (k, Cin, Cout) fp32 (net.py:251-255,264-265) ->
 *   w_fwd (k, cout_pad, [hi cin_pad | lo cin_pad]) bf16: the B operand of the forward GEMM
 *   (K-major) and, read MN-major, of the input-gradient GEMM — one copy serves both. */
int sl_pack_weights(const float* w_keras, void* w_fwd, int k, int Cin, int Cout,
                    int cin_pad, int cout_pad, int prec, void* stream);

/* ---- Conv1D tower (replaces keras.layers.Conv1D(padding="same"), net.py:304-305) ---- */

/* y = act(bias + conv1d_same(x, w)).  x_packed (B,T_in_alloc,..), w_fwd from
 * sl_pack_weights, bias fp32 (Cout).  T_out = ceil(T_in/stride); pad_l per TF
 * SAME rule is computed inside.  T_out_alloc >= T_out (0 = T_out) is the row
 * allocation per utterance of y_packed: a stride-2 consumer wants an even
 * number of rows, the extra row must be (and stays) zero.
 *  act NONE/RELU : y_packed (B,T_out_alloc,[hi cout_pad|lo]) bf16; optional
 *                  relu_mask_out (B,T_out,cout_pad/8) bytes: bit c%8 of byte c/8
 *                  is set iff y[b,t,c] > 0 — the ReLU derivative kept for backward
 *  act SOFTMAX   : probs (B,T_out,Cout) fp32  [net.py:328-330]; optional
 *                  logits (B,T_out,Cout) fp32 (pre-softmax, for parity tests) and
 *                  logp (B,T_out,64) fp32 = log_softmax(log(p+1e-8)), the
 *                  quantity tf.nn.ctc_loss consumes after K.ctc_batch_cost. */
int sl_conv1d_fwd(const void* x_packed, const void* w_fwd, const float* bias,
                  void* y_packed, void* relu_mask_out, float* probs, float* logits,
                  float* logp, int B, int T_in, int T_in_alloc, int T_out_alloc,
                  int Cin, int Cout, int k, int stride, int act, int prec,
                  void* stream);

/* Inverted dropout in front of a Conv1D (keras.layers.Dropout, net.py:301-303), training
 * phase: y = keep ? x/(1-p) : 0 on a packed (B,T_alloc,round64(C)) activation (rows T..T_alloc
 * of y are written as zeros).  The keep decision of an element is a pure function of
 * (seed, b, t, c).  mask_out (B,T,round64(C)/8) bytes = keep bits AND relu_mask_in bits
 * (if given; same dense (B,T,..) layout): with out_scale = 1/(1-p) it is what
 * sl_conv1d_dgrad of the consuming layer applies on its way down. */
int sl_dropout_fwd(const void* x_packed, void* y_packed, const void* relu_mask_in,
                   void* mask_out, int B, int T, int T_alloc, int C, int prec,
                   float p, uint64_t seed, void* stream);

/* dX = out_scale * conv1d_same_backward_input(dY, w) masked by the ReLU of the layer below
 * (TF autodiff of net.py:304-305).  T = frames of dX (the layer's input), dY has
 * ceil(T/stride) frames; stride 2 (striding_conv behind a raw-wave layer, net.py:310-316)
 * runs as one stride-1 problem per output parity.  relu_mask = the relu_mask_out the
 * layer below wrote in its forward pass ((B,T,cin_pad/8) bytes), or NULL for a
 * linear layer below. */
int sl_conv1d_dgrad(const void* dy_packed, const void* w_fwd,
                    const void* relu_mask, void* dx_packed, int B, int T,
                    int Cin, int Cout, int k, int stride, int prec,
                    float out_scale, void* workspace, size_t workspace_bytes,
                    void* stream);
/* Optional fp32 scratch for the split-K variant (long tap loops whose tile count fills the
 * persistent grid badly, e.g. big_conv_1); 0 = not wanted for this shape.  Passing
 * workspace = NULL is always legal and selects the unsplit kernel. */
size_t sl_conv1d_dgrad_workspace_bytes(int B, int T, int Cin, int Cout, int k);

/* dW (k,cout_pad,cin_pad) fp32 += sum_{b,t} x[b,t*s+j-pad_l,ci]*dy[b,t,co];
 * db (Cout) fp32 = sum_{b,t} dy.  dW/db are overwritten (accumulate=0) or
 * accumulated into (accumulate=1).  out_scale multiplies both sums on their way
 * out (1 unless dy carries a loss scale, SL_PREC_FP16). */
int sl_conv1d_wgrad(const void* x_packed, const void* dy_packed, float* dw,
                    float* db, int B, int T_in, int T_in_alloc, int Cin,
                    int Cout, int k, int stride, int prec, int accumulate,
                    float out_scale, void* stream);

/* master-weight layout helpers: Keras (k,Cin,Cout) fp32 <-> internal (k,cout_pad,cin_pad) fp32 */
int sl_weights_keras_to_internal(const float* w_keras, float* w_int, int k,
                                 int Cin, int Cout, int cin_pad, int cout_pad,
                                 void* stream);
int sl_weights_internal_to_keras(const float* w_int, float* w_keras, int k,
                                 int Cin, int Cout, int cin_pad, int cout_pad,
                                 void* stream);
/* internal fp32 master (k,cout_pad,cin_pad) -> packed bf16 w_fwd */
int sl_pack_weights_internal(const float* w_int, void* w_fwd, int k, int cin_pad,
                             int cout_pad, int prec, void* stream);

/* ---- CTC (replaces K.ctc_batch_cost -> tf.nn.ctc_loss, net.py:402-406) ---- */

size_t sl_ctc_workspace_bytes(int B, int T, int L_max);
/* logp (B,T,64) fp32 natural-log probabilities (from sl_conv1d_fwd);
 * labels (B,L_max) int32 padded with -1 (grapheme_enconding.py:25-32);
 * input_len[b] = prediction length P_b (net.py:582), label_len[b];
 * loss (B) fp32 = -log p(label|x).
 * If dlogits_packed != NULL also writes the gradient of
 *   grad_scale * sum_b loss_b  w.r.t. the pre-softmax logits, chained through
 *   log(p+1e-8) and the softmax (probs required), as packed bf16 rows of 64
 *   channels ([hi 64 | lo 64] in SL_PREC_BF16X2), zero for t >= P_b;
 *   optional fp32 copy dlogits_f32 (B,T,V) for tests. */
int sl_ctc_loss(const float* logp, const float* probs, const int32_t* labels,
                const int32_t* input_len, const int32_t* label_len,
                float* loss, void* dlogits_packed, float* dlogits_f32,
                float grad_scale, int B, int T, int V, int L_max, int blank,
                int prec, void* workspace, size_t workspace_bytes,
                void* stream);

/* ---- greedy decode (replaces tf.nn.ctc_greedy_decoder, net.py:453-454) ---- */
/* argmax over V (lowest index wins), drop repeats (merge_repeated) and blanks;
 * out (B,T) int32 padded with -1 (tf.sparse_to_dense default, net.py:436),
 * out_len (B). */
int sl_ctc_greedy_decode(const float* probs, const int32_t* input_len,
                         int32_t* out, int32_t* out_len, int B, int T, int V,
                         int blank, int merge_repeated, void* stream);

/* ---- beam-search decode (replaces tf.nn.ctc_beam_search_decoder, net.py:444-451, stock scorer) ---- */
/* scores (B,T,V) fp32: softmax probabilities (inputs_are_probs = 1: log(p + 1e-8) is taken first, as
 * net.py:430 does) or unnormalised log-scores; every frame is log-softmax-normalised like TF's
 * CTCBeamSearchDecoder::Step.  Prefix beam search of width beam_width (1..128, beam_width * V <= 4096);
 * the top_paths best prefixes of utterance b go to out[b, p, :] (int32, -1 padded to T), out_len[b, p],
 * out_logp[b, p] (log P(prefix | x)), best first.  merge_repeated follows TF: it only drops, from the
 * OUTPUT, a label equal to its predecessor ("A A _ A A" -> [0] with, [0, 0] without; reference
 * test_ctc_decoders.py:38-39) — the reference calls it with merge_repeated = 0 (net.py:441-447).
 * Stock scorer; the language-model scorer is sl_ctc_beam_search_decode_lm below. */
size_t sl_ctc_beam_search_workspace_bytes(int B, int T, int beam_width);
int sl_ctc_beam_search_decode(const float* scores, const int32_t* input_len, int32_t* out,
                              int32_t* out_len, float* out_logp, int B, int T, int V, int blank,
                              int beam_width, int top_paths, int merge_repeated,
                              int inputs_are_probs, void* workspace, size_t workspace_bytes,
                              void* stream);

/* ---- beam-search decode with a word n-gram model inside the search (the reference's KenLM branch:
 * tf.nn.ctc_beam_search_decoder(kenlm_directory_path, kenlm_weight = .8, word_count_weight = 0,
 * valid_word_count_weight = 2.3) of a patched TensorFlow, net.py:420-422,444-451) ---- */
/* The scorer hooks of TF's decoder (ExpandState / GetStateExpansionScore / ExpandStateEnd) run on the
 * device: the letters of the unfinished word walk a vocabulary trie, the space label scores the finished
 * word with an ARPA back-off model, an unfinished word carries the lowest unigram score below its trie node
 * as a look-ahead.  A finished hypothesis collects
 *   weight * log10 P_lm(words </s>) + word_count_weight * #words + valid_word_count_weight * #known words
 * on top of its CTC log-probability (weight = kenlm_weight * ln 10); out_logp holds that sum.  All table
 * pointers are DEVICE pointers owned by the caller (speechless_b200/language_model.py builds them from an
 * ARPA file); the struct itself is passed from host memory.  Semantics: oracle/beam_search_oracle.py
 * (WordLanguageModelScorer; parity with the fork is unpinned — its source is not in the reference tree). */
typedef struct SlWordLm {
  const int32_t* trie_children;  /* (n_trie_nodes, n_labels): child node per symbol, -1 = none; node 0 = root */
  const int32_t* trie_word;      /* (n_trie_nodes): id of the word that ends at the node, -1 = none */
  const float* trie_min_unigram; /* (n_trie_nodes): lowest unigram log10 probability of the words below */
  const int32_t* ngrams;         /* (ngram_mask + 1, 8): open-addressing table {n, id0..id4 (-1 padded),
                                    log10 p, log10 back-off as float bits}; n = 0 marks an empty slot; slot =
                                    FNV-1a over the n ids (32 bit), linear probing */
  uint32_t ngram_mask;           /* table size - 1 (a power of two) */
  int32_t n_labels;              /* = V of the decode call */
  int32_t space_label;           /* the symbol that ends a word */
  int32_t order;                 /* n-gram order, 1..5 */
  int32_t bos_id, eos_id, unk_id; /* word ids of <s>, </s>, <unk>; unk_id is any unused id when has_unk = 0 */
  int32_t has_unk;
  float unknown_log10;           /* unigram score of <unk> (or -100 when the model has none): what a word
                                    outside the model, or a prefix outside the trie, costs */
  float weight;                  /* kenlm_weight * ln 10 */
  float word_count_weight;
  float valid_word_count_weight;
} SlWordLm;
int sl_ctc_beam_search_decode_lm(const float* scores, const int32_t* input_len, int32_t* out,
                                 int32_t* out_len, float* out_logp, int B, int T, int V, int blank,
                                 int beam_width, int top_paths, int merge_repeated,
                                 int inputs_are_probs, const SlWordLm* lm, void* workspace,
                                 size_t workspace_bytes, void* stream);

/* ---- audio front end (replaces the librosa pipeline of labeled_example.py:99-140) ---- */
/* audio (B, audio_stride) fp32 raw samples (sample_counts[b] valid each) ->
 * out[b, t, m] for t < 1 + sample_counts[b]/hop: mel projection (mel_t = transposed Slaney
 * filterbank (257, 128) fp32) of the power level 10 log10 |STFT|^2 floored at -150 dB, with
 * librosa.stft semantics (periodic Hann window, center=True, reflect padding).  Rows beyond
 * an utterance's frame count are left untouched.  Reference defaults only (512 / 128 / 128). */
int sl_spectrogram(const float* audio, const int32_t* sample_counts, const float* mel_t,
                   float* out, int B, int audio_stride, int T_max, int n_fft,
                   int hop_length, int n_mels, void* stream);
/* In place, per utterance: (x - mean) / std over the valid (frame_counts[b] x F) block
 * (labeled_example.py:28-29), zeros beyond it (the batch padding of net.py:583-585).
 * moments_ws: 2*B doubles of scratch. */
int sl_z_normalize(float* x, const int32_t* frame_counts, void* moments_ws, int B,
                   int T_max, int F, void* stream);

/* ---- optimizer (replaces keras.optimizers.Adam, net.py:132,389) ---- */
/* Keras-2 Adam on a flat fp32 buffer: t = step (1-based);
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps). */
int sl_adam_step(float* p, const float* g, float* m, float* v, size_t n,
                 float lr, float beta1, float beta2, float eps, int t,
                 void* stream);

/* Same update fused with the refresh of the bf16 operands: for each of the n_layers kernels
 * placed at floats [w_begin, w_end) of the flat buffer (internal layout, rows of cin_pad),
 * the updated weights are re-emitted into w_fwd[i] (NULL = skip).  The *_host arrays are
 * HOST arrays of length n_layers (<= 16). */
int sl_adam_step_fused(float* p, const float* g, float* m, float* v, size_t n,
                       const size_t* w_begin_host, const size_t* w_end_host,
                       void* const* w_fwd_host, const int* cin_pad_host, int n_layers,
                       int prec, float lr, float beta1, float beta2, float eps, int t,
                       void* stream);

/* ---- data-parallel exchange step (the reference is single-process, so this replaces
 * nothing in it: it is the collective the mean-over-batch objective of net.py:389 needs once the
 * minibatch is sharded by utterance) ---- */
#define SL_COMM_ID_BYTES 128
/* rank 0: writes SL_COMM_ID_BYTES bytes of rendezvous id into id_out (HOST memory); the caller hands
 * them to every rank (any side channel). */
int sl_comm_unique_id(void* id_out_host);
/* every rank, on its own GPU (cudaSetDevice first): joins the communicator of `nranks` processes.
 * max_ctas > 0 bounds the CTAs the collective kernels may occupy (they run next to persistent
 * one-CTA-per-SM tensor-core kernels); 0 = NCCL's default.  *comm_out is an opaque handle. */
int sl_comm_init_rank(void** comm_out, const void* id_host, int nranks, int rank, int max_ctas);
int sl_comm_size(void* comm);
/* in-place SUM all-reduce of count floats at buf (device), stream ordered */
int sl_allreduce_sum(void* comm, float* buf, size_t count, void* stream);
int sl_comm_destroy(void* comm);
int sl_comm_nccl_version(void);

/* Upper bound on the CTAs of the persistent Conv1D kernels launched from now on by this process
 * (0 = every SM of the device).  The data-parallel engine lowers it while a gradient bucket's
 * all-reduce is in flight so that the collective's CTAs find free SMs instead of delaying the last
 * CTAs of a 148-CTA grid. */
int sl_set_sm_limit(int max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* SPEECHLESS_B200_H */
