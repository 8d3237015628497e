// api.cu — the extern "C" surface declared in include/speechless_b200.h
#include "../../include/speechless_b200.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <cstdio>
#include <vector>

#include "conv_umma.h"
#include "tmap.h"

namespace sl {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " at " + file +
                 ":" + std::to_string(line);
  return SL_ERR_CUDA;
}

// launchers implemented in the other translation units
int pack_activation_launch(const float*, void*, int, int, int, int, int, int, cudaStream_t);
int unpack_activation_launch(const void*, float*, int, int, int, int, int, int, cudaStream_t);
int window_activation_launch(const float*, void*, int, int, int, int, int, int, int, int, int, float,
                             unsigned long long, cudaStream_t);
int keras_to_internal_launch(const float*, float*, int, int, int, int, int, cudaStream_t);
int internal_to_keras_launch(const float*, float*, int, int, int, int, int, cudaStream_t);
int pack_weights_internal_launch(const float*, void*, int, int, int, int, cudaStream_t);
int dgrad_finalize_launch(const float*, const void*, void*, size_t, int, int, float, cudaStream_t);
int spectrogram_launch(const float*, const int32_t*, const float*, float*, int, int, int, cudaStream_t);
int z_normalize_launch(float*, const int32_t*, double*, int, int, int, cudaStream_t);
int dropout_launch(const void*, void*, const void*, void*, int, int, int, int, int, float, unsigned long long,
                   cudaStream_t);
int adam_fused_launch(float*, const float*, float*, float*, size_t, const size_t*, const size_t*, void* const*,
                      const int*, int, int, float, float, float, float, int, cudaStream_t);
int adam_launch(float*, const float*, float*, float*, size_t, float, float, float, float, int,
                cudaStream_t);
size_t ctc_workspace_bytes(int B, int T, int L_max);
int ctc_loss_launch(const float*, const float*, const int32_t*, const int32_t*, const int32_t*,
                    float*, void*, float*, float, int, int, int, int, int, int, void*, size_t,
                    cudaStream_t);
int ctc_greedy_launch(const float*, const int32_t*, int32_t*, int32_t*, int, int, int, int, int,
                      cudaStream_t);
size_t beam_search_workspace_bytes(int B, int T, int beam_width);
int beam_search_launch(const float*, const int32_t*, int32_t*, int32_t*, float*, int, int, int, int, int, int, int,
                       int, const SlWordLm*, void*, size_t, cudaStream_t);

static int round64(int c) { return (c + 63) & ~63; }
static int dbg_mode() {
  const char* e = std::getenv("SL_DBG_MODE");  // measurement aid for kernel bring-up; results are garbage
  return e ? std::atoi(e) : 0;
}
static int planes_of(int prec) { return prec_planes(prec); }

// Measurement aid (env SL_TIMELINE=1, tools/selftest perf only): the conv GEMM kernels stamp clock64() at
// their phase boundaries; after every launch the host synchronises and prints, over the CTAs, the median and
// the maximum of each phase in SM cycles.  Never set in production (it serialises every launch).
static long long* timeline_buffer() {
  static long long* buf = nullptr;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = std::getenv("SL_TIMELINE");
    enabled = e != nullptr && std::atoi(e) != 0;
  }
  if (!enabled) return nullptr;
  if (buf == nullptr && cudaMalloc(&buf, 1024 * 32 * sizeof(long long)) != cudaSuccess) return nullptr;
  cudaMemset(buf, 0, 1024 * 32 * sizeof(long long));
  return buf;
}
static void timeline_report(const char* what, const long long* dev) {
  if (dev == nullptr) return;
  static std::vector<long long> h(1024 * 32);
  if (cudaDeviceSynchronize() != cudaSuccess) return;
  cudaMemcpy(h.data(), dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  struct Phase { const char* name; int from, to; };
  const Phase phases[] = {{"entry->setup", 0, 1},      {"setup->pdl", 1, 2},         {"pdl->first operands", 2, 3},
                          {"operands->tile0 issued", 3, 4}, {"tile0->tile1 issued", 4, 5}, {"tile1->tile2 issued", 5, 6},
                          {"tile0 issued->complete", 4, 10}, {"tile0 epilogue", 10, 16},    {"tile1 epilogue", 11, 17},
                          {"tile2 epilogue", 12, 18},   {"tile1 first operands after tile0 issue", 4, 26},
                          {"last store drained->exit sync", 22, 23}, {"entry->exit", 0, 23}, {"dealloc", 23, 24}};
  std::fprintf(stderr, "timeline %s (SM cycles: median / max over CTAs)\n", what);
  for (const Phase& ph : phases) {
    std::vector<long long> d;
    for (int c = 0; c < 1024; ++c) {
      const long long a = h[c * 32 + ph.from], b = h[c * 32 + ph.to];
      if (a != 0 && b != 0) d.push_back(b - a);
    }
    if (d.empty()) continue;
    std::sort(d.begin(), d.end());
    std::fprintf(stderr, "  %-42s %8lld / %8lld  (%zu CTAs)\n", ph.name, d[d.size() / 2], d.back(), d.size());
  }
}
static bool valid_prec(int prec) { return prec == SL_PREC_BF16 || prec == SL_PREC_BF16X2 || prec == SL_PREC_FP16; }

static int g_sm_limit = 0;  // sl_set_sm_limit: CTAs the persistent conv grids may use (0 = all SMs)
static int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& n = cached[dev & 63];
  if (n == 0 && (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)) n = 148;
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

// TF "SAME" padding
static void same_padding(int T, int k, int stride, int* T_out, int* pad_l) {
  *T_out = (T + stride - 1) / stride;
  int total = (*T_out - 1) * stride + k - T;
  if (total < 0) total = 0;
  *pad_l = total / 2;
}

// rank-4 view {C_total, stride, T_alloc/stride, B} of a packed activation
static int make_act_load_map(CUtensorMap* m, const void* base, int c_total, int stride, int T_alloc,
                             int B, int box_rows) {
  const uint64_t row_bytes = static_cast<uint64_t>(c_total) * 2;
  const uint64_t dims[4] = {static_cast<uint64_t>(c_total), static_cast<uint64_t>(stride),
                            static_cast<uint64_t>(T_alloc / stride), static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {row_bytes, row_bytes * stride, row_bytes * T_alloc};
  const uint32_t box[4] = {64, 1, static_cast<uint32_t>(box_rows), 1};
  return make_tmap(m, TMAP_BF16, 4, base, dims, strides, box, true);
}
// rank-3 view {C_total, T, B} of a tensor with T_alloc >= T rows per utterance, every row_step-th row
static int make_act_map3(CUtensorMap* m, const void* base, int c_total, int T, int B, int box_rows,
                         int T_alloc = 0, int row_step = 1) {
  const uint64_t row_bytes = static_cast<uint64_t>(c_total) * 2;
  const uint64_t dims[3] = {static_cast<uint64_t>(c_total), static_cast<uint64_t>(T),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[2] = {row_bytes * row_step, row_bytes * (T_alloc > 0 ? T_alloc : T)};
  const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
  return make_tmap(m, TMAP_BF16, 3, base, dims, strides, box, true);
}
// weights {K_total, rows, taps}
static int make_weight_map(CUtensorMap* m, const void* base, int k_total, int rows, int taps,
                           int box_rows) {
  const uint64_t row_bytes = static_cast<uint64_t>(k_total) * 2;
  const uint64_t dims[3] = {static_cast<uint64_t>(k_total), static_cast<uint64_t>(rows),
                            static_cast<uint64_t>(taps)};
  const uint64_t strides[2] = {row_bytes, row_bytes * rows};
  const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
  return make_tmap(m, TMAP_BF16, 3, base, dims, strides, box, true);
}

// Tail of the persistent conv grid (ConvGemmParams::tail_split / tail_ksplit).  Default: the tiles of the last,
// partial wave are cut into up to 4 narrower tiles (SL_TAIL_SPLIT=0 disables; SL_TAIL_SPLIT_ALL=n splits EVERY
// tile): an M = 128 MMA costs the same ~147 cycles for N = 64 as for N = 256, so this only shortens the
// epilogue of the wave (inner_conv: 42 vs 44 us).  SL_TAIL_KSPLIT=n (opt-in) splits those tiles over the
// contraction instead — the experiment of DESIGN.md §4.1: parity-green, measured slower.
struct TailScratch {
  float* data;
  unsigned* tickets;
  unsigned* done;
};
constexpr int kTailSlots = 148;  // tiles of a partial wave: fewer than the grid
// one zero-filled scratch buffer per (device, stream): launches on one stream are ordered, launches on
// different streams may overlap and must not share partial sums
static int tail_scratch(cudaStream_t stream, int bn, TailScratch* out) {
  struct Entry {
    int dev;
    cudaStream_t stream;
    TailScratch s;
  };
  static Entry table[32];
  static int used = 0;
  static std::mutex mu;
  int dev = 0;
  SL_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used; ++i)
    if (table[i].dev == dev && table[i].stream == stream) {
      *out = table[i].s;
      return 0;
    }
  if (used == 32) {
    out->data = nullptr;  // caller falls back to whole tiles
    return 0;
  }
  (void)bn;
  const size_t data_bytes = static_cast<size_t>(kTailSlots) * 128 * 256 * sizeof(float);
  uint8_t* base = nullptr;
  SL_CUDA(cudaMalloc(&base, data_bytes + 2 * kTailSlots * sizeof(unsigned)));
  SL_CUDA(cudaMemset(base, 0, data_bytes + 2 * kTailSlots * sizeof(unsigned)));
  SL_CUDA(cudaDeviceSynchronize());
  Entry& e = table[used++];
  e.dev = dev;
  e.stream = stream;
  e.s.data = reinterpret_cast<float*>(base);
  e.s.tickets = reinterpret_cast<unsigned*>(base + data_bytes);
  e.s.done = e.s.tickets + kTailSlots;
  *out = e.s;
  return 0;
}
static void plan_tail(int total_tiles, int bn, int chunks, bool allow, int* full_tiles, int* split, int* ksplit,
                      int* per_split) {
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  const int rest = total_tiles % grid;
  *full_tiles = total_tiles;
  *split = 1;
  *ksplit = 1;
  *per_split = chunks;
  const char* e = std::getenv("SL_TAIL_SPLIT");      // max narrow-tile split factor (0/1 = off)
  const char* all = std::getenv("SL_TAIL_SPLIT_ALL");  // measurement aid: split EVERY tile by this factor
  if (allow && all && std::atoi(all) > 1 && bn / std::atoi(all) >= 64) {
    *full_tiles = 0;
    *split = std::atoi(all);
    return;
  }
  if (!allow || rest == 0) return;
  const char* ke = std::getenv("SL_TAIL_KSPLIT");  // max tail K split (default off: measured slower)
  const int max_ksplit = ke ? std::atoi(ke) : 0;
  if (max_ksplit >= 2 && chunks >= 2 && rest <= kTailSlots && bn == 256) {
    // cost of the partial wave in tile times: rounds x the share of the contraction per item, plus the
    // reduce / fold epilogue of a split tile
    double best = 1.0;
    for (int s = 2; s <= max_ksplit && s <= chunks; ++s) {
      const int per = (chunks + s - 1) / s;
      const int items = (chunks + per - 1) / per;  // non-empty splits
      const int rounds = (rest * items + grid - 1) / grid;
      const double cost = rounds * (static_cast<double>(per) / chunks + 0.2);
      if (cost < best - 0.05) {
        best = cost;
        *ksplit = items;
        *per_split = per;
      }
    }
    if (*ksplit > 1) {
      *full_tiles = total_tiles - rest;
      return;
    }
  }
  const int max_split = e ? std::atoi(e) : 4;
  if (max_split < 2) return;
  double best = 1.0;  // cost of the partial wave in units of one full-width tile
  for (int s = 2; s <= max_split && bn / s >= 64; s *= 2) {
    const double cost = static_cast<double>((rest * s + grid - 1) / grid) / s;
    if (cost < best - 1e-9) {
      best = cost;
      *split = s;
    }
  }
  if (*split > 1) *full_tiles = total_tiles - rest;
}
// fills the scratch fields of a plan that chose the tail K split (or reverts it when no scratch is to be had)
static int bind_tail_scratch(ConvGemmParams* p, int bn, cudaStream_t stream) {
  if (p->tail_ksplit <= 1) return 0;
  TailScratch s;
  int rc = tail_scratch(stream, bn, &s);
  if (rc) return rc;
  if (s.data == nullptr) {
    p->tail_ksplit = 1;
    p->full_tiles = p->m_units * p->n_tiles;
    return 0;
  }
  p->tail_scratch = s.data;
  p->tail_tickets = s.tickets;
  p->tail_done = s.done;
  const uint64_t dims[3] = {static_cast<uint64_t>(bn), 128, static_cast<uint64_t>(kTailSlots)};
  const uint64_t strides[2] = {static_cast<uint64_t>(bn) * 4, static_cast<uint64_t>(bn) * 4 * 128};
  const uint32_t box[3] = {32, 32, 1};
  return make_tmap(&p->tmScratch, TMAP_F32, 3, s.data, dims, strides, box, true);
}

// halo mode plan (ConvGemmParams::halo): SL_HALO=0 disables, SL_HALO_BASE selects the descriptor
// base-offset convention (bring-up)
static int make_act_load_map(CUtensorMap* m, const void* base, int c_total, int stride, int T_alloc, int B,
                             int box_rows);
static int plan_halo(ConvGemmParams* p, const void* act, int c_total, int T_alloc, int B, int taps, int stride) {
  // default: on for stride-1 filters of 5 taps and more.  (While the MMA issue loop was bound by its own
  // instruction stream — DESIGN.md §7.1 — the halo only paid for long filters: big_conv_1, k = 32, -4 % bf16,
  // -15 % split-bf16 on its input gradient, and k = 7 lost a little.  With the loop at the tensor pipe's pace
  // the saved operand traffic shows: inner_conv, k = 7, forward 0.040 -> 0.037 ms, input gradient 0.040 -> 0.036.)
  // SL_HALO=0 disables, SL_HALO=1 forces it for every stride-1 layer with more than one tap.
  const char* e = std::getenv("SL_HALO");
  const int mode = e ? std::atoi(e) : -1;
  p->halo = 0;
  if (mode == 0 || stride != 1 || taps < 2 || 128 + taps - 1 > 256) return 0;
  if (mode < 0 && taps < 5) return 0;
  if (mode < 0 && p->ctas == 2) return 0;  // a CTA pair is faster without it (0.905 vs 0.95 ms, big_conv_1 forward)
  p->halo = 1;
  p->halo_rows = 128 + taps - 1;
  // measured on B200: the descriptor start address may simply be advanced by whole 128-byte rows
  // (the swizzle is a function of the absolute smem address); setting the "base offset" field to
  // (address >> 7) & 7 instead gives wrong results.  SL_HALO_BASE=1 reproduces that experiment.
  const char* b = std::getenv("SL_HALO_BASE");
  p->halo_base_mode = b ? std::atoi(b) : 0;
  return make_act_load_map(&p->tmAhalo, act, c_total, 1, T_alloc, B, p->halo_rows);
}

// CTA pairs (ConvGemmParams::ctas, cta_group::2).  Measured on B200 (profiles/README.md): correct, but not
// faster than single CTAs with A-halo reuse — big_conv_1 forward 0.905 ms either way without the halo for
// the pair, 0.95 ms with it; the short inner / striding layers lose 10-15 % — so the pair kernel is
// opt-in: SL_CTA2=1 uses it for every 256-filter tile, SL_CTA2=2 only for layers with >= 2 filter tiles,
// SL_CTA2=3 only for 1-tap layers with >= 2 filter tiles (big_conv_2: the one place it measured faster after
// the issue-loop fix, profiles/r02_cta_pairs_after_fix.log; not validated as a default yet).
static int cta_pairs() {
  const char* e = std::getenv("SL_CTA2");
  return e ? std::atoi(e) : 0;
}
static int grouped_tma() {
  const char* e = std::getenv("SL_GROUPED_TMA");  // 0 disables the one-TMA-per-operand tensor maps
  return e ? std::atoi(e) : 1;
}
// MN-major operands: 64-channel groups as their own tensor dimension so that one TMA box
// {64 ch, rows, n groups} fills the whole [group][row][64 ch] operand tile.
// activation (B, T_alloc, C_total), optionally with the stride-`stride` parity split
static int make_act_group_map(CUtensorMap* m, const void* base, int c_total, int stride, int T_alloc, int B,
                              int box_rows, int box_groups, bool parity_dim) {
  const uint64_t row_bytes = static_cast<uint64_t>(c_total) * 2;
  if (!parity_dim) {
    const uint64_t dims[4] = {64, static_cast<uint64_t>(T_alloc), static_cast<uint64_t>(c_total / 64),
                              static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {row_bytes, 128, row_bytes * T_alloc};
    const uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), static_cast<uint32_t>(box_groups), 1};
    return make_tmap(m, TMAP_BF16, 4, base, dims, strides, box, true);
  }
  const uint64_t dims[5] = {64, static_cast<uint64_t>(stride), static_cast<uint64_t>(T_alloc / stride),
                            static_cast<uint64_t>(c_total / 64), static_cast<uint64_t>(B)};
  const uint64_t strides[4] = {row_bytes, row_bytes * stride, 128, row_bytes * T_alloc};
  const uint32_t box[5] = {64, 1, static_cast<uint32_t>(box_rows), static_cast<uint32_t>(box_groups), 1};
  return make_tmap(m, TMAP_BF16, 5, base, dims, strides, box, true);
}
// weights (taps, rows, K_total): {64, rows, K_total/64, taps}
static int make_weight_group_map(CUtensorMap* m, const void* base, int k_total, int rows, int taps,
                                 int box_rows, int box_groups) {
  const uint64_t row_bytes = static_cast<uint64_t>(k_total) * 2;
  const uint64_t dims[4] = {64, static_cast<uint64_t>(rows), static_cast<uint64_t>(k_total / 64),
                            static_cast<uint64_t>(taps)};
  const uint64_t strides[3] = {row_bytes, 128, row_bytes * rows};
  const uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), static_cast<uint32_t>(box_groups), 1};
  return make_tmap(m, TMAP_BF16, 4, base, dims, strides, box, true);
}

}  // namespace sl

using namespace sl;

extern "C" {

int sl_version(void) { return 100; }

int sl_last_error(char* buf, size_t n) {
  if (buf == nullptr || n == 0) return SL_ERR_INVALID;
  std::strncpy(buf, g_last_error.c_str(), n - 1);
  buf[n - 1] = '\0';
  return SL_OK;
}

int sl_set_sm_limit(int max_ctas) {
  SL_REQUIRE(max_ctas >= 0, "limit must be >= 0");
  g_sm_limit = max_ctas;
  return SL_OK;
}

int sl_sync_check(void) {
  SL_CUDA(cudaDeviceSynchronize());
  SL_CUDA(cudaGetLastError());
  return SL_OK;
}

int sl_pack_activation(const float* x, void* x_packed, int B, int T, int C, int T_alloc, int c_pad,
                       int prec, void* stream) {
  SL_REQUIRE(x && x_packed, "null pointer");
  SL_REQUIRE(B > 0 && T > 0 && C > 0 && T_alloc >= T, "bad shape");
  SL_REQUIRE(c_pad % 64 == 0 && c_pad >= C, "c_pad must be a multiple of 64 and >= C");
  SL_REQUIRE(valid_prec(prec), "bad precision");
  return pack_activation_launch(x, x_packed, B, T, C, T_alloc, c_pad, prec,
                                static_cast<cudaStream_t>(stream));
}

int sl_window_activation(const float* x, void* x_windowed, int B, int T, int C, int k, int stride, int c_pad,
                         int prec, float drop_p, unsigned long long seed, void* stream) {
  SL_REQUIRE(x && x_windowed, "null pointer");
  SL_REQUIRE(B > 0 && T > 0 && C > 0 && k > 0 && stride > 0, "bad shape");
  SL_REQUIRE(c_pad % 64 == 0 && c_pad >= k * C, "c_pad must be a multiple of 64 and >= k*C");
  SL_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout rate must be in [0, 1)");
  int T_out, pad_l;
  same_padding(T, k, stride, &T_out, &pad_l);
  SL_REQUIRE(valid_prec(prec), "bad precision");
  return window_activation_launch(x, x_windowed, B, T, C, k, stride, T_out, pad_l, c_pad, prec, drop_p,
                                  seed, static_cast<cudaStream_t>(stream));
}

int sl_unpack_activation(const void* x_packed, float* x, int B, int T, int C, int T_alloc, int c_pad,
                         int prec, void* stream) {
  SL_REQUIRE(x && x_packed, "null pointer");
  SL_REQUIRE(B > 0 && T > 0 && C > 0 && T_alloc >= T && c_pad >= C, "bad shape");
  SL_REQUIRE(valid_prec(prec), "bad precision");
  return unpack_activation_launch(x_packed, x, B, T, C, T_alloc, c_pad, prec,
                                  static_cast<cudaStream_t>(stream));
}

int sl_weights_keras_to_internal(const float* w_keras, float* w_int, int k, int Cin, int Cout,
                                 int cin_pad, int cout_pad, void* stream) {
  SL_REQUIRE(w_keras && w_int, "null pointer");
  SL_REQUIRE(cin_pad >= Cin && cout_pad >= Cout && k > 0, "bad shape");
  return keras_to_internal_launch(w_keras, w_int, k, Cin, Cout, cin_pad, cout_pad,
                                  static_cast<cudaStream_t>(stream));
}
int sl_weights_internal_to_keras(const float* w_int, float* w_keras, int k, int Cin, int Cout,
                                 int cin_pad, int cout_pad, void* stream) {
  SL_REQUIRE(w_keras && w_int, "null pointer");
  SL_REQUIRE(cin_pad >= Cin && cout_pad >= Cout && k > 0, "bad shape");
  return internal_to_keras_launch(w_int, w_keras, k, Cin, Cout, cin_pad, cout_pad,
                                  static_cast<cudaStream_t>(stream));
}
int sl_pack_weights_internal(const float* w_int, void* w_fwd, int k, int cin_pad, int cout_pad, int prec,
                             void* stream) {
  SL_REQUIRE(w_int && w_fwd, "null pointer");
  SL_REQUIRE(cin_pad % 64 == 0 && cout_pad % 64 == 0 && k > 0, "bad shape");
  SL_REQUIRE(valid_prec(prec), "bad precision");
  return pack_weights_internal_launch(w_int, w_fwd, k, cin_pad, cout_pad, prec,
                                      static_cast<cudaStream_t>(stream));
}

int sl_pack_weights(const float* w_keras, void* w_fwd, int k, int Cin, int Cout, int cin_pad, int cout_pad,
                    int prec, void* stream) {
  // convenience path (tests, weight loading): via a temporary internal master copy
  SL_REQUIRE(w_keras && w_fwd, "null pointer");
  SL_REQUIRE(valid_prec(prec), "bad precision");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* tmp = nullptr;
  const size_t bytes = static_cast<size_t>(k) * cin_pad * cout_pad * sizeof(float);
  SL_CUDA(cudaMallocAsync(&tmp, bytes, s));
  int rc = keras_to_internal_launch(w_keras, tmp, k, Cin, Cout, cin_pad, cout_pad, s);
  if (rc == 0) rc = pack_weights_internal_launch(tmp, w_fwd, k, cin_pad, cout_pad, prec, s);
  cudaFreeAsync(tmp, s);
  return rc;
}

int sl_conv1d_fwd(const void* x_packed, const void* w_fwd, const float* bias, void* y_packed,
                  void* relu_mask_out, float* probs, float* logits, float* logp, int B, int T_in,
                  int T_in_alloc, int T_out_alloc, int Cin, int Cout, int k, int stride, int act, int prec,
                  void* stream) {
  SL_REQUIRE(x_packed && w_fwd, "null pointer");
  SL_REQUIRE(B > 0 && T_in > 0 && Cin > 0 && Cout > 0 && k > 0, "bad shape");
  SL_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
  SL_REQUIRE(T_in_alloc >= T_in && T_in_alloc % stride == 0, "T_in_alloc must cover T_in and divide by stride");
  SL_REQUIRE(valid_prec(prec), "bad precision");
  const int planes = planes_of(prec);
  const int cin_pad = round64(Cin), cout_pad = round64(Cout);
  int T_out, pad_l;
  same_padding(T_in, k, stride, &T_out, &pad_l);
  if (T_out_alloc <= 0) T_out_alloc = T_out;
  SL_REQUIRE(T_out_alloc >= T_out, "T_out_alloc must cover the output frames");

  ConvGemmParams p;
  std::memset(&p, 0, sizeof(p));
  int bn = cout_pad >= 256 ? 256 : cout_pad;
  SL_REQUIRE(cout_pad % bn == 0 && (bn == 64 || bn == 128 || bn == 256), "unsupported filter count");
  p.ctas = (bn == 256 && act != SL_ACT_SOFTMAX && dbg_mode() == 0 &&
            (cta_pairs() == 1 || (cta_pairs() == 2 && cout_pad / bn >= 2) ||
             (cta_pairs() == 3 && cout_pad / bn >= 2 && k == 1))) ? 2 : 1;
  int rc = make_act_load_map(&p.tmA, x_packed, planes * cin_pad, stride, T_in_alloc, B, 128);
  if (rc) return rc;
  rc = make_weight_map(&p.tmB, w_fwd, planes * cin_pad, cout_pad, k, bn / p.ctas);
  if (rc) return rc;
  rc = plan_halo(&p, x_packed, planes * cin_pad, T_in_alloc, B, k, stride);
  if (rc) return rc;
  p.B = B;
  p.T_out = T_out;
  p.m_tiles_per_utt = (T_out + 127) / 128;
  p.m_units = (B * p.m_tiles_per_utt + p.ctas - 1) / p.ctas;
  p.n_tiles = cout_pad / bn;
  if (p.ctas == 2) {
    p.full_tiles = p.m_units * p.n_tiles;
    p.tail_split = 1;
  } else {
    plan_tail(B * p.m_tiles_per_utt * p.n_tiles, bn, cin_pad / 64, act != SL_ACT_SOFTMAX, &p.full_tiles, &p.tail_split,
              &p.tail_ksplit, &p.tail_per);
    rc = bind_tail_scratch(&p, bn, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
  }
  if (p.tail_split > 1) {
    rc = make_weight_map(&p.tmBtail, w_fwd, planes * cin_pad, cout_pad, k, bn / p.tail_split);
    if (rc) return rc;
  }
  p.taps = k;
  p.chunks = cin_pad / 64;
  p.terms = planes == 2 ? 3 : 1;
  p.fp16 = prec_fp16(prec);
  p.a_lo_off = cin_pad;
  p.b_lo_off = cin_pad;
  p.stride = stride;
  p.pad_l = pad_l;
  p.w_tap0 = 0;
  p.w_tap_step = 1;
  p.mask_T = T_out;
  p.out_t_scale = 1;
  p.dbg_mode = dbg_mode();
  p.out_scale = 1.0f;
  p.bias = bias;
  p.n_valid = Cout;
  int epi = EPI_PACKED;
  if (act == SL_ACT_SOFTMAX) {
    SL_REQUIRE(probs != nullptr, "softmax epilogue needs the probs output");
    SL_REQUIRE(cout_pad == 64, "softmax epilogue supports up to 64 symbols");
    epi = EPI_SOFTMAX;
    {
      const char* e = std::getenv("SL_REVERSE_OUTPUT");  // 0: first-to-last tile order (A/B runs)
      p.reverse_order = e ? std::atoi(e) : 1;
    }
    p.probs = probs;
    p.logits = logits;
    p.logp = logp;
    p.V = Cout;
  } else {
    SL_REQUIRE(act == SL_ACT_NONE || act == SL_ACT_RELU, "bad activation");
    SL_REQUIRE(y_packed != nullptr, "null output");
    p.relu = act == SL_ACT_RELU;
    p.y_planes = planes;
    p.y_lo_off = cout_pad;
    p.mask_bits_out = static_cast<uint8_t*>(relu_mask_out);
    p.mask_row_bytes = cout_pad / 8;
    rc = make_act_map3(&p.tmY, y_packed, planes * cout_pad, T_out, B, 32, T_out_alloc);
    if (rc) return rc;
  }
  p.timeline = timeline_buffer();
  rc = conv_gemm_launch(p, bn, epi, false, num_sms(), static_cast<cudaStream_t>(stream));
  timeline_report("conv forward", p.timeline);
  return rc;
}

// split-K plan of the input-gradient GEMM: worthwhile when the tile count leaves the last
// wave of the persistent grid mostly empty and every tile is a long loop over many taps
static int plan_dgrad_ksplit(int B, int T, int cin_pad, int cout_pad, int k) {
  const char* e = std::getenv("SL_DGRAD_KSPLIT");
  if (e) return std::atoi(e) > 1 ? std::atoi(e) : 1;
  const int bn = cin_pad >= 256 ? 256 : cin_pad;
  if (bn < 128 || k < 8) return 1;
  const int tiles = B * ((T + 127) / 128) * (cin_pad / bn);
  const int sms = num_sms();
  if (tiles < sms) return 1;
  const long long ksteps = static_cast<long long>(k) * (cout_pad / 64);
  if (ksteps < 256) return 1;
  int best = 1;
  double best_cost = static_cast<double>((tiles + sms - 1) / sms);
  for (int ks = 2; ks <= 8 && ks <= (cout_pad / 64) / 4; ks *= 2) {  // (the split runs over 64-channel chunks)
    // each split item also pays a fixed epilogue + pipeline ramp (~3 % of a 1024-step tile)
    const double cost = static_cast<double>((tiles * ks + sms - 1) / sms) / ks * (1.0 + 0.01 * ks);
    if (cost < best_cost * 0.95) {
      best_cost = cost;
      best = ks;
    }
  }
  return best;
}

size_t sl_conv1d_dgrad_workspace_bytes(int B, int T, int Cin, int Cout, int k) {
  const int cin_pad = round64(Cin), cout_pad = round64(Cout);
  if (plan_dgrad_ksplit(B, T, cin_pad, cout_pad, k) <= 1) return 0;
  return static_cast<size_t>(B) * T * cin_pad * sizeof(float);
}

// one launch of the input-gradient GEMM: the output rows u = out_scale_t * v + out_off (v = 0..rows-1)
// of dX, filter taps w_tap0, w_tap0 + w_step, ... (n_taps of them), dY frames shifted by i - a_pad
static int dgrad_launch_rows(const void* dy_packed, const void* w_fwd, const void* relu_mask, void* dx_packed, int B,
                             int T, int T_dy, int cin_pad, int cout_pad, int Cin, int k, int prec, float out_scale,
                             int rows, int row_step, int row_off, int n_taps, int w_tap0, int w_step, int a_pad,
                             void* workspace, size_t workspace_bytes, bool allow_ksplit, cudaStream_t s) {
  ConvGemmParams p;
  std::memset(&p, 0, sizeof(p));
  const int planes = planes_of(prec);
  const int bn = cin_pad >= 256 ? 256 : cin_pad;
  SL_REQUIRE(cin_pad % bn == 0 && (bn == 64 || bn == 128 || bn == 256), "unsupported channel count");
  int rc = make_act_load_map(&p.tmA, dy_packed, planes * cout_pad, 1, T_dy, B, 128);
  if (rc) return rc;
  // B[n = ci][k = co] comes straight from the forward layout (k, cout_pad, [hi|lo] cin_pad):
  // boxes of 64 co rows x 64 ci, consumed MN-major
  p.b_grouped = grouped_tma();
  p.ctas = (bn == 256 && p.b_grouped && dbg_mode() == 0 &&
            (cta_pairs() == 1 || (cta_pairs() == 2 && cin_pad / bn >= 2 && cout_pad >= 512) ||
             (cta_pairs() == 3 && cin_pad / bn >= 2 && cout_pad >= 512 && n_taps == 1))) ? 2 : 1;
  rc = p.b_grouped ? make_weight_group_map(&p.tmB, w_fwd, planes * cin_pad, cout_pad, k, 64, bn / 64 / p.ctas)
                   : make_weight_map(&p.tmB, w_fwd, planes * cin_pad, cout_pad, k, 64);
  if (rc) return rc;
  const char* dx_rows = static_cast<const char*>(dx_packed) + static_cast<size_t>(row_off) * planes * cin_pad * 2;
  rc = make_act_map3(&p.tmY, dx_rows, planes * cin_pad, rows, B, 32, T, row_step);
  if (rc) return rc;
  rc = plan_halo(&p, dy_packed, planes * cout_pad, T_dy, B, n_taps, 1);
  if (rc) return rc;
  p.B = B;
  p.T_out = rows;
  p.m_tiles_per_utt = (rows + 127) / 128;
  p.m_units = (B * p.m_tiles_per_utt + p.ctas - 1) / p.ctas;
  p.n_tiles = cin_pad / bn;
  if (p.ctas == 2) {
    p.full_tiles = p.m_units * p.n_tiles;
    p.tail_split = 1;
  } else {
    plan_tail(B * p.m_tiles_per_utt * p.n_tiles, bn, cout_pad / 64, p.b_grouped != 0, &p.full_tiles, &p.tail_split,
              &p.tail_ksplit, &p.tail_per);
    rc = bind_tail_scratch(&p, bn, s);
    if (rc) return rc;
  }
  if (p.tail_split > 1) {
    rc = make_weight_group_map(&p.tmBtail, w_fwd, planes * cin_pad, cout_pad, k, 64, bn / p.tail_split / 64);
    if (rc) return rc;
  }
  p.taps = n_taps;
  p.chunks = cout_pad / 64;
  p.terms = planes == 2 ? 3 : 1;
  p.fp16 = prec_fp16(prec);
  p.a_lo_off = cout_pad;
  p.b_lo_off = cin_pad;
  p.stride = 1;
  p.pad_l = a_pad;
  p.w_tap0 = w_tap0;
  p.w_tap_step = w_step;
  p.dbg_mode = dbg_mode();
  p.bias = nullptr;
  p.n_valid = Cin;
  p.relu = 0;
  p.out_scale = out_scale;
  p.y_planes = planes;
  p.y_lo_off = cin_pad;
  p.mask_bits_in = static_cast<const uint8_t*>(relu_mask);
  p.mask_row_bytes = cin_pad / 8;
  p.mask_T = T;
  p.out_t_scale = row_step;
  p.out_t_off = row_off;
  p.ksplit = 1;
  const int ksplit = allow_ksplit ? plan_dgrad_ksplit(B, T, cin_pad, cout_pad, k) : 1;
  const size_t need = static_cast<size_t>(B) * T * cin_pad * sizeof(float);
  if (ksplit > 1 && p.b_grouped && workspace != nullptr && workspace_bytes >= need) {
    // split K over tap ranges: fp32 partial sums meet in `workspace` (TMA reduce-add), then one
    // elementwise pass applies the ReLU mask and packs to bf16
    SL_CUDA(cudaMemsetAsync(workspace, 0, need, s));
    const uint64_t dims[3] = {static_cast<uint64_t>(cin_pad), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(cin_pad) * 4, static_cast<uint64_t>(cin_pad) * 4 * T};
    const uint32_t box[3] = {32, 32, 1};
    rc = make_tmap(&p.tmY, TMAP_F32, 3, workspace, dims, strides, box, true);
    if (rc) return rc;
    p.ksplit = ksplit;
    p.full_tiles = p.m_units * p.n_tiles;
    p.tail_split = 1;
    p.tail_ksplit = 1;
    p.mask_bits_in = nullptr;
    rc = conv_gemm_launch(p, bn, EPI_F32, true, num_sms(), s);
    if (rc) return rc;
    return dgrad_finalize_launch(static_cast<const float*>(workspace), relu_mask, dx_packed,
                                 static_cast<size_t>(B) * T, cin_pad, prec, out_scale, s);
  }
  p.timeline = timeline_buffer();
  rc = conv_gemm_launch(p, bn, EPI_PACKED, true, num_sms(), s);
  timeline_report("conv input gradient", p.timeline);
  return rc;
}

int sl_conv1d_dgrad(const void* dy_packed, const void* w_fwd, const void* relu_mask, void* dx_packed,
                    int B, int T, int Cin, int Cout, int k, int stride, int prec, float out_scale,
                    void* workspace, size_t workspace_bytes, void* stream) {
  SL_REQUIRE(dy_packed && w_fwd && dx_packed, "null pointer");
  SL_REQUIRE(B > 0 && T > 0 && Cin > 0 && Cout > 0 && k > 0, "bad shape");
  SL_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
  SL_REQUIRE(valid_prec(prec), "bad precision");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int planes = planes_of(prec);
  const int cin_pad = round64(Cin), cout_pad = round64(Cout);
  int T_out, pad_l;
  same_padding(T, k, stride, &T_out, &pad_l);
  if (stride == 1) {
    // dX[u] = sum_j dY[u + pad_l - j] W[j]  (SURVEY.md A.1): taps walked backwards
    return dgrad_launch_rows(dy_packed, w_fwd, relu_mask, dx_packed, B, T, T_out, cin_pad, cout_pad, Cin, k, prec,
                             out_scale, T, 1, 0, k, k - 1, -1, k - 1 - pad_l, workspace, workspace_bytes, true, s);
  }
  // stride 2: y[v'] = sum_j x[2 v' + j - pad_l] W[j], so the rows u = 2 v + r of dX only see the taps
  // j = jmax_r, jmax_r - 2, ... with (r + pad_l - j) even: dX[2v + r] = sum_i dY[v + c_r + i] W[jmax_r - 2 i],
  // c_r = (r + pad_l - jmax_r) / 2 — one stride-1 problem per output parity, written to every other row
  for (int r = 0; r < 2; ++r) {
    const int rows = (T - r + 1) / 2;
    if (rows <= 0) continue;
    int jmax = k - 1;
    if (((r + pad_l - jmax) & 1) != 0) --jmax;
    if (jmax < 0) {  // k == 1: this parity receives no gradient at all
      const size_t row_bytes = static_cast<size_t>(planes) * cin_pad * 2;
      for (int b = 0; b < B; ++b)
        SL_CUDA(cudaMemset2DAsync(static_cast<char*>(dx_packed) + (static_cast<size_t>(b) * T + r) * row_bytes,
                                  2 * row_bytes, 0, row_bytes, static_cast<size_t>(rows), s));
      continue;
    }
    const int n_taps = jmax / 2 + 1;
    const int c_r = (r + pad_l - jmax) / 2;  // exact: the numerator is even
    const int rc = dgrad_launch_rows(dy_packed, w_fwd, relu_mask, dx_packed, B, T, T_out, cin_pad, cout_pad, Cin, k,
                                     prec, out_scale, rows, 2, r, n_taps, jmax, -2, -c_r, nullptr, 0, false, s);
    if (rc) return rc;
  }
  return 0;
}

int sl_conv1d_wgrad(const void* x_packed, const void* dy_packed, float* dw, float* db, int B, int T_in,
                    int T_in_alloc, int Cin, int Cout, int k, int stride, int prec, int accumulate,
                    float out_scale, void* stream) {
  SL_REQUIRE(x_packed && dy_packed && dw, "null pointer");
  SL_REQUIRE(B > 0 && T_in > 0 && Cin > 0 && Cout > 0 && k > 0, "bad shape");
  SL_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
  SL_REQUIRE(T_in_alloc >= T_in && T_in_alloc % stride == 0, "T_in_alloc must cover T_in and divide by stride");
  SL_REQUIRE(valid_prec(prec), "bad precision");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int planes = planes_of(prec);
  const int cin_pad = round64(Cin), cout_pad = round64(Cout);
  int T_out, pad_l;
  same_padding(T_in, k, stride, &T_out, &pad_l);

  WgradParams p;
  std::memset(&p, 0, sizeof(p));
  const int bn = cin_pad >= 256 ? 256 : cin_pad;
  SL_REQUIRE(cin_pad % bn == 0 && (bn == 64 || bn == 128 || bn == 256), "unsupported channel count");
  p.grouped = grouped_tma();
  // 128 input channels (striding_conv on 128 mel bins): pair adjacent taps into N = 256 tiles
  const char* pair_env = std::getenv("SL_WGRAD_TAP_PAIR");  // tuning aid: 0 disables
  const bool tap_pair = bn == 128 && k >= 2 && p.grouped && !(pair_env && std::atoi(pair_env) == 0);
  p.tap_step = tap_pair ? 2 : 1;
  p.tap_units = tap_pair ? (k + 1) / 2 : k;
  const int bn_tile = tap_pair ? 256 : bn;  // columns of the accumulator tile
  int rc = p.grouped ? make_act_group_map(&p.tmDY, dy_packed, planes * cout_pad, 1, T_out, B, 64, 2, false)
                     : make_act_map3(&p.tmDY, dy_packed, planes * cout_pad, T_out, B, 64);
  if (rc) return rc;
  rc = p.grouped ? make_act_group_map(&p.tmX, x_packed, planes * cin_pad, stride, T_in_alloc, B, 64, bn / 64, true)
                 : make_act_load_map(&p.tmX, x_packed, planes * cin_pad, stride, T_in_alloc, B, 64);
  if (rc) return rc;
  p.dw = dw;
  p.B = B;
  p.T_out = T_out;
  p.taps = k;
  p.m_tiles = (cout_pad + 127) / 128;
  p.n_tiles = cin_pad / bn;
  p.tchunks = (T_out + 63) / 64;
  p.terms = planes == 2 ? 3 : 1;
  p.fp16 = prec_fp16(prec);
  p.out_scale = out_scale;
  p.dy_lo_off = cout_pad;
  p.x_lo_off = cin_pad;
  p.stride = stride;
  p.pad_l = pad_l;
  p.cout_pad = cout_pad;
  p.cin_pad = cin_pad;
  p.dy_c_total = planes * cout_pad;
  p.dbg_mode = dbg_mode();
  {
    // fp32 view of dW for the TMA (reduce-)store epilogue
    const uint64_t dims[3] = {static_cast<uint64_t>(cin_pad), static_cast<uint64_t>(cout_pad),
                              static_cast<uint64_t>(k)};
    const uint64_t strides[2] = {static_cast<uint64_t>(cin_pad) * 4, static_cast<uint64_t>(cin_pad) * cout_pad * 4};
    const uint32_t box[3] = {32, 32, 1};
    rc = make_tmap(&p.tmDW, TMAP_F32, 3, dw, dims, strides, box, true);
    if (rc) return rc;
  }
  // K split: minimise  waves x (pipeline steps per unit + epilogue)  over the split count
  const int base_units = p.tap_units * p.m_tiles * p.n_tiles;
  const int k_total = B * p.tchunks;
  int ksplit = 1;
  {
    const int sms = num_sms();
    const double step_cycles = 512.0 * bn_tile / 256.0 * p.terms;  // 4 MMAs of 128 x bn x 16 per stage
    const double epilogue_cycles = 1500.0 * bn_tile / 32.0 / 8.0 + 2000.0;
    double best = 1e300;
    for (int ks = 1; ks <= k_total && ks <= 256; ++ks) {
      const int per = (k_total + ks - 1) / ks;
      if (per * (ks - 1) >= k_total) continue;  // would leave an empty split
      const int waves = (base_units * ks + sms - 1) / sms;
      const double cost = waves * (per * step_cycles + epilogue_cycles);
      if (cost < best * 0.98) {  // prefer fewer splits (less reduction traffic) on near ties
        best = cost;
        ksplit = ks;
      }
    }
  }
  if (const char* e = std::getenv("SL_WGRAD_KSPLIT")) {  // tuning / test aid: force the K split
    const int forced = std::atoi(e);
    if (forced >= 1 && forced <= k_total && ((k_total + forced - 1) / forced) * (forced - 1) < k_total) ksplit = forced;
  }
  p.ksplit = ksplit;
  p.use_atomics = (ksplit > 1 || accumulate) ? 1 : 0;
  const size_t dw_bytes = static_cast<size_t>(k) * cout_pad * cin_pad * sizeof(float);
  if (p.use_atomics && !accumulate) SL_CUDA(cudaMemsetAsync(dw, 0, dw_bytes, s));
  // bias gradient: fused into the wgrad kernel (column sums of the staged dY tiles)
  p.db = db;
  p.n_filters = Cout;
  if (db != nullptr && !accumulate) SL_CUDA(cudaMemsetAsync(db, 0, static_cast<size_t>(Cout) * sizeof(float), s));
  return wgrad_launch(p, bn_tile, num_sms(), s);
}

size_t sl_ctc_workspace_bytes(int B, int T, int L_max) {
  return ctc_workspace_bytes(B, T, L_max < 1 ? 1 : L_max);
}

int sl_ctc_loss(const float* logp, const float* probs, const int32_t* labels, const int32_t* input_len,
                const int32_t* label_len, float* loss, void* dlogits_packed, float* dlogits_f32,
                float grad_scale, int B, int T, int V, int L_max, int blank, int prec, void* workspace,
                size_t workspace_bytes, void* stream) {
  SL_REQUIRE(logp && labels && input_len && label_len && loss && workspace, "null pointer");
  SL_REQUIRE(B > 0 && T > 0, "bad shape");
  return ctc_loss_launch(logp, probs, labels, input_len, label_len, loss, dlogits_packed, dlogits_f32,
                         grad_scale, B, T, V, L_max, blank, prec, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

int sl_ctc_greedy_decode(const float* probs, const int32_t* input_len, int32_t* out, int32_t* out_len,
                         int B, int T, int V, int blank, int merge_repeated, void* stream) {
  SL_REQUIRE(probs && input_len && out && out_len, "null pointer");
  SL_REQUIRE(B > 0 && T > 0 && V > 0, "bad shape");
  return ctc_greedy_launch(probs, input_len, out, out_len, B, T, V, blank, merge_repeated,
                           static_cast<cudaStream_t>(stream));
}

size_t sl_ctc_beam_search_workspace_bytes(int B, int T, int beam_width) {
  return beam_search_workspace_bytes(B, T, beam_width < 1 ? 1 : beam_width);
}

int sl_ctc_beam_search_decode(const float* scores, const int32_t* input_len, int32_t* out, int32_t* out_len,
                              float* out_logp, int B, int T, int V, int blank, int beam_width, int top_paths,
                              int merge_repeated, int inputs_are_probs, void* workspace, size_t workspace_bytes,
                              void* stream) {
  SL_REQUIRE(scores && input_len && out && out_len && out_logp && workspace, "null pointer");
  SL_REQUIRE(B > 0 && T > 0 && V > 0, "bad shape");
  return beam_search_launch(scores, input_len, out, out_len, out_logp, B, T, V, blank, beam_width, top_paths,
                            merge_repeated, inputs_are_probs, nullptr, workspace, workspace_bytes,
                            static_cast<cudaStream_t>(stream));
}

int sl_ctc_beam_search_decode_lm(const float* scores, const int32_t* input_len, int32_t* out, int32_t* out_len,
                                 float* out_logp, int B, int T, int V, int blank, int beam_width, int top_paths,
                                 int merge_repeated, int inputs_are_probs, const SlWordLm* lm, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  SL_REQUIRE(scores && input_len && out && out_len && out_logp && workspace && lm, "null pointer");
  SL_REQUIRE(B > 0 && T > 0 && V > 0, "bad shape");
  return beam_search_launch(scores, input_len, out, out_len, out_logp, B, T, V, blank, beam_width, top_paths,
                            merge_repeated, inputs_are_probs, lm, workspace, workspace_bytes,
                            static_cast<cudaStream_t>(stream));
}

int sl_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                 float beta2, float eps, int t, void* stream) {
  SL_REQUIRE(p && g && m && v, "null pointer");
  SL_REQUIRE(t >= 1, "Adam step counter is 1-based");
  SL_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
              reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
             "Adam buffers must be 16-byte aligned");
  if (n == 0) return SL_OK;
  return adam_launch(p, g, m, v, n, lr, beta1, beta2, eps, t, static_cast<cudaStream_t>(stream));
}

int sl_spectrogram(const float* audio, const int32_t* sample_counts, const float* mel_t, float* out, int B,
                   int audio_stride, int T_max, int n_fft, int hop_length, int n_mels, void* stream) {
  SL_REQUIRE(audio && sample_counts && mel_t && out, "null pointer");
  SL_REQUIRE(B > 0 && audio_stride > 0 && T_max > 0, "bad shape");
  SL_REQUIRE(n_fft == 512 && hop_length == 128 && n_mels == 128,
             "the front end is built for the reference defaults: n_fft 512, hop 128, 128 mel bins");
  return spectrogram_launch(audio, sample_counts, mel_t, out, B, audio_stride, T_max,
                            static_cast<cudaStream_t>(stream));
}

int sl_z_normalize(float* x, const int32_t* frame_counts, void* moments_ws, int B, int T_max, int F, void* stream) {
  SL_REQUIRE(x && frame_counts && moments_ws, "null pointer");
  SL_REQUIRE(B > 0 && T_max > 0 && F > 0, "bad shape");
  SL_REQUIRE(reinterpret_cast<uintptr_t>(moments_ws) % 8 == 0, "moments workspace must be 8-byte aligned");
  return z_normalize_launch(x, frame_counts, static_cast<double*>(moments_ws), B, T_max, F,
                            static_cast<cudaStream_t>(stream));
}

int sl_dropout_fwd(const void* x_packed, void* y_packed, const void* relu_mask_in, void* mask_out, int B, int T,
                   int T_alloc, int C, int prec, float p, uint64_t seed, void* stream) {
  SL_REQUIRE(x_packed && y_packed && mask_out, "null pointer");
  SL_REQUIRE(B > 0 && T > 0 && C > 0 && T_alloc >= T, "bad shape");
  SL_REQUIRE(p >= 0.0f && p < 1.0f, "dropout rate must be in [0, 1)");
  SL_REQUIRE(valid_prec(prec), "bad precision");
  return dropout_launch(x_packed, y_packed, relu_mask_in, mask_out, B, T, T_alloc, round64(C), prec, p,
                        seed, static_cast<cudaStream_t>(stream));
}

int sl_adam_step_fused(float* p, const float* g, float* m, float* v, size_t n, const size_t* w_begin_host,
                       const size_t* w_end_host, void* const* w_fwd_host, const int* cin_pad_host, int n_layers,
                       int prec, float lr, float beta1, float beta2, float eps, int t, void* stream) {
  SL_REQUIRE(p && g && m && v && w_begin_host && w_end_host && w_fwd_host && cin_pad_host, "null pointer");
  SL_REQUIRE(t >= 1, "Adam step counter is 1-based");
  SL_REQUIRE(n % 4 == 0, "flat buffer length must be a multiple of 4");
  SL_REQUIRE(n_layers >= 0 && n_layers <= 16, "at most 16 layers");
  SL_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
              reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
             "Adam buffers must be 16-byte aligned");
  for (int i = 0; i < n_layers; ++i)
    SL_REQUIRE(w_begin_host[i] % 4 == 0 && w_end_host[i] % 4 == 0 && cin_pad_host[i] % 64 == 0 &&
                   w_end_host[i] <= n && w_begin_host[i] <= w_end_host[i],
               "bad layer placement");
  if (n == 0) return SL_OK;
  return adam_fused_launch(p, g, m, v, n, w_begin_host, w_end_host, w_fwd_host, cin_pad_host, n_layers,
                           prec, lr, beta1, beta2, eps, t, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
