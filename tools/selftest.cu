// selftest.cu — standalone GPU bring-up harness for libspeechless_b200.so.
// Test infrastructure only: runs every kernel through the C-ABI on small shapes and
// compares with straightforward double-precision CPU loops written here (no torch, no
// python: a fresh GPU box spends its minutes on kernels, not imports).
//   usage: tools/selftest [filter-substring]
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "speechless_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA failure %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)
#define SLCK(x)                                                   \
  do {                                                            \
    int rc_ = (x);                                                \
    if (rc_ != 0) {                                               \
      char buf[1024];                                             \
      sl_last_error(buf, sizeof buf);                             \
      printf("  sl error %d: %s  (%s:%d)\n", rc_, buf, __FILE__, __LINE__); \
      return false;                                               \
    }                                                             \
  } while (0)

static int round64(int c) { return (c + 63) & ~63; }
static float bf16r(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

template <class T>
struct Dev {
  T* p = nullptr;
  size_t n = 0;
  explicit Dev(size_t n_) : n(n_) {
    CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    CK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
  }
  ~Dev() { cudaFree(p); }
  void up(const std::vector<T>& h) { CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
  std::vector<T> down() const {
    std::vector<T> h(n);
    CK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
  }
};

static std::mt19937 rng(1234);
static std::vector<float> randn(size_t n, float scale = 1.f) {
  std::normal_distribution<float> d(0.f, 1.f);
  std::vector<float> v(n);
  for (auto& x : v) x = d(rng) * scale;
  return v;
}

struct Cmp {
  double max_err = 0, max_ref = 0;
  size_t bad = 0, first_bad = 0;
};
static Cmp compare(const std::vector<float>& got, const std::vector<double>& want, double rel_tol) {
  Cmp c;
  for (size_t i = 0; i < want.size(); ++i) c.max_ref = std::max(c.max_ref, std::fabs(want[i]));
  const double tol = rel_tol * std::max(c.max_ref, 1e-30);
  for (size_t i = 0; i < want.size(); ++i) {
    const double e = std::fabs(got[i] - want[i]);
    if (!(e <= tol)) {  // catches NaN
      if (c.bad == 0) c.first_bad = i;
      ++c.bad;
    }
    if (e > c.max_err || std::isnan(e)) c.max_err = e;
  }
  return c;
}
static bool report(const char* what, const Cmp& c, const std::vector<float>& got,
                   const std::vector<double>& want) {
  printf("  %-26s max_err %.3e  max_ref %.3e  rel %.3e  bad %zu/%zu", what, c.max_err, c.max_ref,
         c.max_err / std::max(c.max_ref, 1e-30), c.bad, want.size());
  if (c.bad) printf("  first_bad@%zu got %.6g want %.6g", c.first_bad, got[c.first_bad], want[c.first_bad]);
  printf("\n");
  return c.bad == 0;
}

static void same_pad(int T, int k, int s, int* T_out, int* pad_l) {
  *T_out = (T + s - 1) / s;
  int total = std::max((*T_out - 1) * s + k - T, 0);
  *pad_l = total / 2;
}

// ---------------- CPU references (double) ----------------
// x (B,T,Cin), w keras (k,Cin,Cout), bias (Cout) -> y (B,T_out,Cout) pre-activation
static std::vector<double> cpu_conv(const std::vector<float>& x, const std::vector<float>& w,
                                    const std::vector<float>& bias, int B, int T, int Cin, int Cout,
                                    int k, int s) {
  int T_out, pad_l;
  same_pad(T, k, s, &T_out, &pad_l);
  std::vector<double> y(static_cast<size_t>(B) * T_out * Cout);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T_out; ++t) {
      std::vector<double> acc(Cout);
      for (int co = 0; co < Cout; ++co) acc[co] = bias.empty() ? 0.0 : bias[co];
      for (int j = 0; j < k; ++j) {
        const int ti = t * s + j - pad_l;
        if (ti < 0 || ti >= T) continue;
        const float* xr = &x[(static_cast<size_t>(b) * T + ti) * Cin];
        for (int ci = 0; ci < Cin; ++ci) {
          const double xv = xr[ci];
          if (xv == 0.0) continue;
          const float* wr = &w[(static_cast<size_t>(j) * Cin + ci) * Cout];
          for (int co = 0; co < Cout; ++co) acc[co] += xv * wr[co];
        }
      }
      for (int co = 0; co < Cout; ++co) y[(static_cast<size_t>(b) * T_out + t) * Cout + co] = acc[co];
    }
  return y;
}
// dX (B,T,Cin) = sum_j sum_co dY[b, u+pad_l-j, co] W[j,ci,co]  (stride 1)
static std::vector<double> cpu_dgrad(const std::vector<float>& dy, const std::vector<float>& w, int B,
                                     int T, int Cin, int Cout, int k) {
  int T_out, pad_l;
  same_pad(T, k, 1, &T_out, &pad_l);
  std::vector<double> dx(static_cast<size_t>(B) * T * Cin);
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int u = 0; u < T; ++u) {
      std::vector<double> acc(Cin, 0.0);
      for (int j = 0; j < k; ++j) {
        const int t = u + pad_l - j;
        if (t < 0 || t >= T) continue;
        const float* dr = &dy[(static_cast<size_t>(b) * T + t) * Cout];
        for (int ci = 0; ci < Cin; ++ci) {
          const float* wr = &w[(static_cast<size_t>(j) * Cin + ci) * Cout];
          double a = 0;
          for (int co = 0; co < Cout; ++co) a += static_cast<double>(dr[co]) * wr[co];
          acc[ci] += a;
        }
      }
      for (int ci = 0; ci < Cin; ++ci) dx[(static_cast<size_t>(b) * T + u) * Cin + ci] = acc[ci];
    }
  return dx;
}
// dW keras (k,Cin,Cout), db (Cout)
static void cpu_wgrad(const std::vector<float>& x, const std::vector<float>& dy, int B, int T, int Cin,
                      int Cout, int k, int s, std::vector<double>* dw, std::vector<double>* db) {
  int T_out, pad_l;
  same_pad(T, k, s, &T_out, &pad_l);
  dw->assign(static_cast<size_t>(k) * Cin * Cout, 0.0);
  db->assign(Cout, 0.0);
#pragma omp parallel for collapse(2)
  for (int j = 0; j < k; ++j)
    for (int ci = 0; ci < Cin; ++ci) {
      std::vector<double> acc(Cout, 0.0);
      for (int b = 0; b < B; ++b)
        for (int t = 0; t < T_out; ++t) {
          const int ti = t * s + j - pad_l;
          if (ti < 0 || ti >= T) continue;
          const double xv = x[(static_cast<size_t>(b) * T + ti) * Cin + ci];
          if (xv == 0.0) continue;
          const float* dr = &dy[(static_cast<size_t>(b) * T_out + t) * Cout];
          for (int co = 0; co < Cout; ++co) acc[co] += xv * dr[co];
        }
      for (int co = 0; co < Cout; ++co) (*dw)[(static_cast<size_t>(j) * Cin + ci) * Cout + co] = acc[co];
    }
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T_out; ++t)
      for (int co = 0; co < Cout; ++co) (*db)[co] += dy[(static_cast<size_t>(b) * T_out + t) * Cout + co];
}

static std::vector<float> round_bf16(std::vector<float> v) {
  for (auto& x : v) x = bf16r(x);
  return v;
}
static std::vector<float> round_fp16(std::vector<float> v) {
  for (auto& x : v) x = __half2float(__float2half_rn(x));
  return v;
}
// operands as the kernels see them: bf16 (prec 1) / fp16 (prec 3) round each operand once, the split-bf16
// mode (prec 2) carries ~16 mantissa bits
static std::vector<float> round_for(int prec, const std::vector<float>& v) {
  return prec == 1 ? round_bf16(v) : (prec == 3 ? round_fp16(v) : v);
}
static int planes_of(int prec) { return prec == 2 ? 2 : 1; }
static const char* prec_name(int prec) { return prec == 1 ? "bf16" : (prec == 2 ? "bf16x2" : "fp16"); }

// ---------------- tests ----------------
static bool test_pack() {
  const int B = 3, T = 37, C = 250, T_alloc = 38, cp = 256;
  auto x = randn(static_cast<size_t>(B) * T * C);
  Dev<float> dx(x.size()), dback(x.size());
  dx.up(x);
  bool ok = true;
  for (int prec = 1; prec <= 3; ++prec) {
    Dev<uint16_t> dp(static_cast<size_t>(B) * T_alloc * cp * planes_of(prec));
    SLCK(sl_pack_activation(dx.p, dp.p, B, T, C, T_alloc, cp, prec, nullptr));
    SLCK(sl_unpack_activation(dp.p, dback.p, B, T, C, T_alloc, cp, prec, nullptr));
    SLCK(sl_sync_check());
    auto got = dback.down();
    std::vector<double> want(x.begin(), x.end());
    auto c = compare(got, want, prec == 1 ? 4e-3 : (prec == 3 ? 5e-4 : 2e-5));
    ok &= report((std::string("pack/unpack ") + prec_name(prec)).c_str(), c, got, want);
  }
  return ok;
}

struct ConvCase {
  const char* name;
  int B, T, Cin, Cout, k, s, act;
};

static bool run_conv_fwd(const ConvCase& cc, int prec) {
  const int B = cc.B, T = cc.T, Cin = cc.Cin, Cout = cc.Cout, k = cc.k, s = cc.s;
  int T_out, pad_l;
  same_pad(T, k, s, &T_out, &pad_l);
  const int T_alloc = (T + s - 1) / s * s;
  const int cip = round64(Cin), cop = round64(Cout);
  auto x = randn(static_cast<size_t>(B) * T * Cin);
  const float wscale = 1.f / std::sqrt(static_cast<float>(k) * Cin);
  auto w = randn(static_cast<size_t>(k) * Cin * Cout, wscale);
  auto bias = randn(Cout, 0.5f);
  Dev<float> dx(x.size()), dw(w.size()), dbias(bias.size());
  dx.up(x);
  dw.up(w);
  dbias.up(bias);
  Dev<uint16_t> xp(static_cast<size_t>(B) * T_alloc * cip * planes_of(prec));
  Dev<uint16_t> wf(static_cast<size_t>(k) * cop * cip * planes_of(prec));
  SLCK(sl_pack_activation(dx.p, xp.p, B, T, Cin, T_alloc, cip, prec, nullptr));
  SLCK(sl_pack_weights(dw.p, wf.p, k, Cin, Cout, cip, cop, prec, nullptr));

  // reference: bf16 mode multiplies bf16-rounded operands exactly; bf16x2 ~ fp32 operands
  std::vector<double> ref = cpu_conv(round_for(prec, x), round_for(prec, w), bias, B, T, Cin, Cout, k, s);
  bool ok = true;
  char label[128];
  if (cc.act == SL_ACT_SOFTMAX) {
    Dev<float> probs(static_cast<size_t>(B) * T_out * Cout), logits(probs.n), logp(static_cast<size_t>(B) * T_out * 64);
    SLCK(sl_conv1d_fwd(xp.p, wf.p, dbias.p, nullptr, nullptr, probs.p, logits.p, logp.p, B, T, T_alloc, 0, Cin, Cout,
                       k, s, SL_ACT_SOFTMAX, prec, nullptr));
    SLCK(sl_sync_check());
    auto gl = logits.down(), gp = probs.down(), glp = logp.down();
    std::vector<double> rp(ref.size()), rlp(ref.size());
    for (size_t r = 0; r < ref.size() / Cout; ++r) {
      double m = -1e300, sum = 0, sume = 0;
      for (int v = 0; v < Cout; ++v) m = std::max(m, ref[r * Cout + v]);
      for (int v = 0; v < Cout; ++v) sum += std::exp(ref[r * Cout + v] - m);
      for (int v = 0; v < Cout; ++v) {
        rp[r * Cout + v] = std::exp(ref[r * Cout + v] - m) / sum;
        sume += rp[r * Cout + v] + 1e-8;
      }
      for (int v = 0; v < Cout; ++v) rlp[r * Cout + v] = std::log(rp[r * Cout + v] + 1e-8) - std::log(sume);
    }
    std::vector<float> glp_c(ref.size());
    for (size_t r = 0; r < ref.size() / Cout; ++r)
      for (int v = 0; v < Cout; ++v) glp_c[r * Cout + v] = glp[r * 64 + v];
    const double tol = prec == 2 ? 1e-4 : 2e-3;
    snprintf(label, sizeof label, "%s logits p%d", cc.name, prec);
    ok &= report(label, compare(gl, ref, tol), gl, ref);
    snprintf(label, sizeof label, "%s probs p%d", cc.name, prec);
    ok &= report(label, compare(gp, rp, tol * 5), gp, rp);
    snprintf(label, sizeof label, "%s logp p%d", cc.name, prec);
    ok &= report(label, compare(glp_c, rlp, tol), glp_c, rlp);
  } else {
    if (cc.act == SL_ACT_RELU)
      for (auto& v : ref) v = std::max(v, 0.0);
    Dev<uint16_t> yp(static_cast<size_t>(B) * T_out * cop * planes_of(prec));
    Dev<uint8_t> mask(static_cast<size_t>(B) * T_out * cop / 8);
    Dev<float> y(static_cast<size_t>(B) * T_out * Cout);
    SLCK(sl_conv1d_fwd(xp.p, wf.p, dbias.p, yp.p, mask.p, nullptr, nullptr, nullptr, B, T, T_alloc, 0, Cin, Cout, k, s,
                       cc.act, prec, nullptr));
    SLCK(sl_unpack_activation(yp.p, y.p, B, T_out, Cout, T_out, cop, prec, nullptr));
    SLCK(sl_sync_check());
    auto got = y.down();
    snprintf(label, sizeof label, "%s p%d", cc.name, prec);
    // bf16 output rounding: 2^-9 relative; split: 2^-17
    ok &= report(label, compare(got, ref, prec == 1 ? 6e-3 : (prec == 3 ? 1e-3 : 5e-5)), got, ref);
    if (cc.act == SL_ACT_RELU) {
      // the ReLU bitmask must agree with the sign of the stored activation
      auto mb = mask.down();
      size_t bad = 0;
      for (size_t r = 0; r < static_cast<size_t>(B) * T_out; ++r)
        for (int c = 0; c < Cout; ++c) {
          const bool bit = (mb[r * (cop / 8) + c / 8] >> (c % 8)) & 1;
          const bool pos = got[r * Cout + c] > 0.f;
          // values that round to zero in bf16 may legitimately differ
          if (bit != pos && std::fabs(ref[r * Cout + c]) > 1e-30) ++bad;
        }
      printf("  %-26s relu bitmask mismatches %zu\n", label, bad);
      ok &= bad == 0;
    }
  }
  return ok;
}

static bool run_dgrad(const ConvCase& cc, int prec) {
  const int B = cc.B, T = cc.T, Cin = cc.Cin, Cout = cc.Cout, k = cc.k;
  const int cip = round64(Cin), cop = round64(Cout);
  auto dy = randn(static_cast<size_t>(B) * T * Cout);
  auto w = randn(static_cast<size_t>(k) * Cin * Cout, 1.f / std::sqrt(static_cast<float>(k) * Cout));
  auto xs = randn(static_cast<size_t>(B) * T * Cin);  // saved activation: relu'd
  for (auto& v : xs) v = std::max(v, 0.f);
  Dev<float> ddy(dy.size()), dw(w.size()), dxs(xs.size());
  ddy.up(dy);
  dw.up(w);
  dxs.up(xs);
  Dev<uint16_t> dyp(static_cast<size_t>(B) * T * cop * planes_of(prec)), dxp(static_cast<size_t>(B) * T * cip * planes_of(prec));
  Dev<uint16_t> wd(static_cast<size_t>(k) * cip * cop * planes_of(prec));
  Dev<float> dxo(static_cast<size_t>(B) * T * Cin);
  SLCK(sl_pack_activation(ddy.p, dyp.p, B, T, Cout, T, cop, prec, nullptr));
  std::vector<uint8_t> hmask(static_cast<size_t>(B) * T * cip / 8, 0);
  for (size_t r = 0; r < static_cast<size_t>(B) * T; ++r)
    for (int c = 0; c < Cin; ++c)
      if (xs[r * Cin + c] > 0.f) hmask[r * (cip / 8) + c / 8] |= static_cast<uint8_t>(1u << (c % 8));
  Dev<uint8_t> dmask(hmask.size());
  dmask.up(hmask);
  SLCK(sl_pack_weights(dw.p, wd.p, k, Cin, Cout, cip, cop, prec, nullptr));
  const size_t wsb = sl_conv1d_dgrad_workspace_bytes(B, T, Cin, Cout, k);
  Dev<uint8_t> dws(wsb);
  if (wsb) printf("  (split-K dgrad, %zu byte scratch)\n", wsb);
  SLCK(sl_conv1d_dgrad(dyp.p, wd.p, cc.act == SL_ACT_RELU ? dmask.p : nullptr, dxp.p, B, T, Cin, Cout, k, 1, prec,
                       1.0f, wsb ? dws.p : nullptr, wsb, nullptr));
  SLCK(sl_unpack_activation(dxp.p, dxo.p, B, T, Cin, T, cip, prec, nullptr));
  SLCK(sl_sync_check());
  auto ref = cpu_dgrad(round_for(prec, dy), round_for(prec, w), B, T, Cin, Cout, k);
  if (cc.act == SL_ACT_RELU)
    for (size_t i = 0; i < ref.size(); ++i)
      if (!(xs[i] > 0.f)) ref[i] = 0.0;
  auto got = dxo.down();
  char label[128];
  snprintf(label, sizeof label, "%s dgrad p%d", cc.name, prec);
  // tcgen05 accumulates in fp32 with truncation: the error grows with the number of
  // accumulation steps (measured ~2e-4 of max at K = 32 taps x 2048 channels x 3 terms)
  const double tol2 = static_cast<double>(k) * cop > 8192 ? 5e-4 : 5e-5;
  return report(label, compare(got, ref, prec == 1 ? 6e-3 : (prec == 3 ? 1e-3 : tol2)), got, ref);
}

static bool run_wgrad(const ConvCase& cc, int prec) {
  const int B = cc.B, T = cc.T, Cin = cc.Cin, Cout = cc.Cout, k = cc.k, s = cc.s;
  int T_out, pad_l;
  same_pad(T, k, s, &T_out, &pad_l);
  const int T_alloc = (T + s - 1) / s * s;
  const int cip = round64(Cin), cop = round64(Cout);
  auto x = randn(static_cast<size_t>(B) * T * Cin);
  auto dy = randn(static_cast<size_t>(B) * T_out * Cout);
  Dev<float> dx(x.size()), ddy(dy.size());
  dx.up(x);
  ddy.up(dy);
  Dev<uint16_t> xp(static_cast<size_t>(B) * T_alloc * cip * planes_of(prec)), dyp(static_cast<size_t>(B) * T_out * cop * planes_of(prec));
  Dev<float> dwi(static_cast<size_t>(k) * cop * cip), dwk(static_cast<size_t>(k) * Cin * Cout), db(Cout);
  SLCK(sl_pack_activation(dx.p, xp.p, B, T, Cin, T_alloc, cip, prec, nullptr));
  SLCK(sl_pack_activation(ddy.p, dyp.p, B, T_out, Cout, T_out, cop, prec, nullptr));
  SLCK(sl_conv1d_wgrad(xp.p, dyp.p, dwi.p, db.p, B, T, T_alloc, Cin, Cout, k, s, prec, 0, 1.0f, nullptr));
  SLCK(sl_weights_internal_to_keras(dwi.p, dwk.p, k, Cin, Cout, cip, cop, nullptr));
  SLCK(sl_sync_check());
  std::vector<double> rdw, rdb;
  cpu_wgrad(round_for(prec, x), round_for(prec, dy), B, T, Cin, Cout, k, s, &rdw, &rdb);
  auto gdw = dwk.down(), gdb = db.down();
  char label[128];
  snprintf(label, sizeof label, "%s wgrad p%d", cc.name, prec);
  bool ok = report(label, compare(gdw, rdw, prec == 2 ? 5e-5 : 1e-4), gdw, rdw);
  snprintf(label, sizeof label, "%s bgrad p%d", cc.name, prec);
  ok &= report(label, compare(gdb, rdb, prec == 2 ? 5e-5 : 1e-4), gdb, rdb);
  return ok;
}

// ---- CTC reference (double) ----
static double lse(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const double m = std::max(a, b);
  return m + std::log(std::exp(a - m) + std::exp(b - m));
}
static bool test_ctc(int B, int T, int V, int L_lo, int L_hi, const char* name, float logit_scale = 2.0f,
                     bool tight = false) {
  const int blank = V - 1;
  std::uniform_int_distribution<int> dl(L_lo, L_hi), dv(0, V - 2);
  std::vector<int> label_len(B), input_len(B);
  int L_max = 1;
  for (int b = 0; b < B; ++b) {
    label_len[b] = dl(rng);
    L_max = std::max(L_max, label_len[b]);
  }
  std::vector<int> labels(static_cast<size_t>(B) * L_max, -1);
  for (int b = 0; b < B; ++b) {
    int rep = 0;
    for (int i = 0; i < label_len[b]; ++i) {
      labels[b * L_max + i] = (i > 0 && (rng() % 4 == 0)) ? labels[b * L_max + i - 1] : dv(rng);
      if (i > 0 && labels[b * L_max + i] == labels[b * L_max + i - 1]) ++rep;
    }
    const int need = label_len[b] + rep;
    input_len[b] = std::min(T, std::max(need, T - static_cast<int>(rng() % (T / 3 + 1))));
    if (tight) input_len[b] = std::min(T, need + static_cast<int>(rng() % 2));  // (almost) a single alignment
    if (need > T) printf("  (case %d infeasible: need %d > T %d)\n", b, need, T);
  }
  // logits -> probs (double), logp
  auto z = randn(static_cast<size_t>(B) * T * V, logit_scale);
  std::vector<double> p(z.size()), lp(z.size());
  std::vector<float> pf(z.size()), lpf(static_cast<size_t>(B) * T * 64, -INFINITY);
  for (size_t r = 0; r < z.size() / V; ++r) {
    double m = -1e300, sum = 0, sume = 0;
    for (int v = 0; v < V; ++v) m = std::max(m, static_cast<double>(z[r * V + v]));
    for (int v = 0; v < V; ++v) sum += std::exp(z[r * V + v] - m);
    for (int v = 0; v < V; ++v) {
      p[r * V + v] = std::exp(z[r * V + v] - m) / sum;
      pf[r * V + v] = static_cast<float>(p[r * V + v]);
    }
    // the GPU consumes fp32 probs; follow the same rounding for the reference chain
    for (int v = 0; v < V; ++v) sume += static_cast<double>(pf[r * V + v]) + 1e-8;
    for (int v = 0; v < V; ++v) {
      lp[r * V + v] = std::log(static_cast<double>(pf[r * V + v]) + 1e-8) - std::log(sume);
      lpf[r * 64 + v] = static_cast<float>(lp[r * V + v]);
    }
  }
  // reference loss + grad wrt logits (scale 1/B)
  const double scale = 1.0 / B;
  std::vector<double> rloss(B), rdz(z.size(), 0.0);
  for (int b = 0; b < B; ++b) {
    const int L = label_len[b], P = input_len[b], S = 2 * L + 1;
    std::vector<int> e(S, blank);
    for (int i = 0; i < L; ++i) e[2 * i + 1] = labels[b * L_max + i];
    std::vector<double> al(static_cast<size_t>(P) * S, -INFINITY), be(al);
    auto LP = [&](int t, int v) { return lp[(static_cast<size_t>(b) * T + t) * V + v]; };
    al[0] = LP(0, blank);
    if (S > 1) al[1] = LP(0, e[1]);
    for (int t = 1; t < P; ++t)
      for (int s = 0; s < S; ++s) {
        double a = al[(t - 1) * S + s];
        if (s >= 1) a = lse(a, al[(t - 1) * S + s - 1]);
        if (s >= 2 && e[s] != blank && e[s] != e[s - 2]) a = lse(a, al[(t - 1) * S + s - 2]);
        al[t * S + s] = a + LP(t, e[s]);
      }
    be[(P - 1) * S + S - 1] = LP(P - 1, blank);
    if (S > 1) be[(P - 1) * S + S - 2] = LP(P - 1, e[S - 2]);
    for (int t = P - 2; t >= 0; --t)
      for (int s = 0; s < S; ++s) {
        double a = be[(t + 1) * S + s];
        if (s + 1 < S) a = lse(a, be[(t + 1) * S + s + 1]);
        if (s + 2 < S && e[s + 2] != blank && e[s + 2] != e[s]) a = lse(a, be[(t + 1) * S + s + 2]);
        be[t * S + s] = a + LP(t, e[s]);
      }
    double ll = al[(P - 1) * S + S - 1];
    if (S > 1) ll = lse(ll, al[(P - 1) * S + S - 2]);
    rloss[b] = -ll;
    for (int t = 0; t < P; ++t) {
      std::vector<double> occ(V, 0.0), dLdp(V);
      for (int s = 0; s < S; ++s) occ[e[s]] += std::exp(al[t * S + s] + be[t * S + s] - LP(t, e[s]) - ll);
      double dot = 0;
      for (int v = 0; v < V; ++v) {
        const double pv = pf[(static_cast<size_t>(b) * T + t) * V + v];
        dLdp[v] = (std::exp(LP(t, v)) - occ[v]) / (pv + 1e-8);
        dot += pv * dLdp[v];
      }
      for (int v = 0; v < V; ++v) {
        const double pv = pf[(static_cast<size_t>(b) * T + t) * V + v];
        rdz[(static_cast<size_t>(b) * T + t) * V + v] = pv * (dLdp[v] - dot) * scale;
      }
    }
  }
  Dev<float> dlp(lpf.size()), dp(pf.size()), dloss(B), ddz(z.size());
  Dev<int> dlab(labels.size()), dil(B), dll(B);
  Dev<uint16_t> dzp(static_cast<size_t>(B) * T * 128);
  dlp.up(lpf);
  dp.up(pf);
  dlab.up(labels);
  dil.up(input_len);
  dll.up(label_len);
  const size_t wsb = sl_ctc_workspace_bytes(B, T, L_max);
  Dev<uint8_t> ws(wsb);
  SLCK(sl_ctc_loss(dlp.p, dp.p, dlab.p, dil.p, dll.p, dloss.p, dzp.p, ddz.p, static_cast<float>(scale), B, T, V,
                   L_max, blank, SL_PREC_BF16X2, ws.p, wsb, nullptr));
  Dev<float> dzu(static_cast<size_t>(B) * T * V);
  SLCK(sl_unpack_activation(dzp.p, dzu.p, B, T, V, T, 64, SL_PREC_BF16X2, nullptr));
  SLCK(sl_sync_check());
  {  // timing of the loss + gradient launch pair (CUDA events, 10 iterations after the run above)
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 10; ++it)
      SLCK(sl_ctc_loss(dlp.p, dp.p, dlab.p, dil.p, dll.p, dloss.p, dzp.p, nullptr, static_cast<float>(scale), B, T, V,
                       L_max, blank, SL_PREC_BF16, ws.p, wsb, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("  %-26s loss + gradient: %.4f ms per call (B=%d T=%d L_max=%d)\n", name, ms / 10, B, T, L_max);
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 10; ++it)
      SLCK(sl_ctc_loss(dlp.p, dp.p, dlab.p, dil.p, dll.p, dloss.p, nullptr, nullptr, static_cast<float>(scale), B, T, V,
                       L_max, blank, SL_PREC_BF16, ws.p, wsb, nullptr));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("  %-26s lattices only:   %.4f ms per call\n", name, ms / 10);
  }
  auto gl = dloss.down(), gdz = ddz.down(), gdzu = dzu.down();
  char label[128];
  bool ok = true;
  // per-utterance relative loss error
  double worst = 0;
  for (int b = 0; b < B; ++b) worst = std::max(worst, std::fabs(gl[b] - rloss[b]) / std::fabs(rloss[b]));
  printf("  %-26s worst rel loss err %.3e (loss[0] gpu %.6f ref %.6f)\n", name, worst, gl[0], rloss[0]);
  ok &= worst < 1e-5;
  // alpha/beta are fp32 log-space values of magnitude ~loss (as in TF's float CTC): the
  // occupancy exp(alpha+beta-...) inherits ~ulp(loss)*sqrt(T) of noise, so the gradient
  // tolerance scales with the loss magnitude.
  double max_loss = 0;
  for (int b = 0; b < B; ++b) max_loss = std::max(max_loss, rloss[b]);
  // (...and with the square root of the number of frames: 626 frames is the bench shape)
  const double gtol = 1e-4 + 3e-6 * max_loss * std::max(1.0, std::sqrt(T / 626.0));
  snprintf(label, sizeof label, "%s dlogits", name);
  ok &= report(label, compare(gdz, rdz, gtol), gdz, rdz);
  snprintf(label, sizeof label, "%s dlogits packed", name);
  ok &= report(label, compare(gdzu, rdz, gtol), gdzu, rdz);

  // greedy decode vs CPU
  Dev<int> dout(static_cast<size_t>(B) * T), dolen(B);
  SLCK(sl_ctc_greedy_decode(dp.p, dil.p, dout.p, dolen.p, B, T, V, blank, 1, nullptr));
  SLCK(sl_sync_check());
  auto gout = dout.down(), golen = dolen.down();
  size_t bad = 0;
  for (int b = 0; b < B; ++b) {
    std::vector<int> want;
    int prev = -1;
    for (int t = 0; t < input_len[b]; ++t) {
      int c = 0;
      for (int v = 1; v < V; ++v)
        if (pf[(static_cast<size_t>(b) * T + t) * V + v] > pf[(static_cast<size_t>(b) * T + t) * V + c]) c = v;
      if (c != blank && c != prev) want.push_back(c);
      prev = c;
    }
    if (golen[b] != static_cast<int>(want.size())) ++bad;
    for (int i = 0; i < T; ++i) {
      const int w_ = i < static_cast<int>(want.size()) ? want[i] : -1;
      if (gout[static_cast<size_t>(b) * T + i] != w_) ++bad;
    }
  }
  printf("  %-26s greedy decode mismatches %zu\n", name, bad);
  ok &= bad == 0;
  return ok;
}

static bool test_greedy_golden() {
  // reference test_ctc_decoders.py:22-24,38-41: "A A blank A A", V=2, blank=1
  std::vector<float> p = {1, 0, 1, 0, 0, 1, 1, 0, 1, 0};
  std::vector<int> il = {5};
  Dev<float> dp(p.size());
  Dev<int> dil(1), dout(5), dol(1);
  dp.up(p);
  dil.up(il);
  bool ok = true;
  for (int merge = 1; merge >= 0; --merge) {
    SLCK(sl_ctc_greedy_decode(dp.p, dil.p, dout.p, dol.p, 1, 5, 2, 1, merge, nullptr));
    SLCK(sl_sync_check());
    auto o = dout.down();
    auto n = dol.down();
    std::vector<int> want = merge ? std::vector<int>{0, 0, -1, -1, -1} : std::vector<int>{0, 0, 0, 0, -1};
    const bool good = o == want && n[0] == (merge ? 2 : 4);
    printf("  greedy golden merge=%d: %s [%d %d %d %d %d] len %d\n", merge, good ? "ok" : "MISMATCH", o[0], o[1],
           o[2], o[3], o[4], n[0]);
    ok &= good;
  }
  return ok;
}

static bool test_adam() {
  const size_t n = 100003;
  auto p = randn(n), g = randn(n, 0.1f);
  std::vector<float> m(n, 0.f), v(n, 0.f);
  Dev<float> dp(n + 1), dg(n + 1), dm(n + 1), dv(n + 1);
  dp.up(p);
  dg.up(g);
  std::vector<double> rp(p.begin(), p.end()), rm(n, 0.0), rv(n, 0.0);
  const double lr = 1e-4, b1 = 0.9, b2 = 0.999, eps = 1e-8;
  for (int t = 1; t <= 3; ++t) {
    SLCK(sl_adam_step(dp.p, dg.p, dm.p, dv.p, n, 1e-4f, 0.9f, 0.999f, 1e-8f, t, nullptr));
    const double lr_t = lr * std::sqrt(1 - std::pow(b2, t)) / (1 - std::pow(b1, t));
    for (size_t i = 0; i < n; ++i) {
      rm[i] = b1 * rm[i] + (1 - b1) * g[i];
      rv[i] = b2 * rv[i] + (1 - b2) * static_cast<double>(g[i]) * g[i];
      rp[i] -= lr_t * rm[i] / (std::sqrt(rv[i]) + eps);
    }
  }
  SLCK(sl_sync_check());
  auto got = dp.down();
  got.resize(n);
  // compare the update, not the parameter: (p - p0)
  std::vector<float> gu(n);
  std::vector<double> ru(n);
  for (size_t i = 0; i < n; ++i) {
    gu[i] = got[i] - p[i];
    ru[i] = rp[i] - p[i];
  }
  bool ok = report("adam 3 steps (update)", compare(gu, ru, 2e-3), gu, ru);
  // fused variant: same update + bf16 re-emission of two "layers" (rows of 64 / 128 channels)
  const size_t n2 = 64 * 64 + 16 + 32 * 128;
  auto p2 = randn(n2), g2 = randn(n2, 0.1f);
  Dev<float> fp(n2), fg(n2), fm(n2), fv(n2);
  fp.up(p2);
  fg.up(g2);
  Dev<uint16_t> w0(64 * 64 * 2), w1(32 * 128 * 2);
  const size_t begins[2] = {0, 64 * 64 + 16}, ends[2] = {64 * 64, n2};
  void* targets[2] = {w0.p, w1.p};
  const int cin_pads[2] = {64, 128};
  SLCK(sl_adam_step_fused(fp.p, fg.p, fm.p, fv.p, n2, begins, ends, targets, cin_pads, 2, SL_PREC_BF16X2, 1e-4f,
                          0.9f, 0.999f, 1e-8f, 1, nullptr));
  Dev<float> back0(64 * 64), back1(32 * 128);
  SLCK(sl_unpack_activation(w0.p, back0.p, 1, 64, 64, 64, 64, SL_PREC_BF16X2, nullptr));
  SLCK(sl_unpack_activation(w1.p, back1.p, 1, 32, 128, 32, 128, SL_PREC_BF16X2, nullptr));
  SLCK(sl_sync_check());
  auto newp = fp.down();
  auto bk0 = back0.down(), bk1 = back1.down();
  std::vector<double> want0(newp.begin(), newp.begin() + 64 * 64), want1(newp.begin() + 64 * 64 + 16, newp.end());
  ok &= report("adam fused: w_fwd layer 0", compare(bk0, want0, 2e-5), bk0, want0);
  ok &= report("adam fused: w_fwd layer 1", compare(bk1, want1, 2e-5), bk1, want1);
  std::vector<double> wantp(n2);
  const double lr1 = 1e-4 * std::sqrt(1 - 0.999) / (1 - 0.9);
  for (size_t i = 0; i < n2; ++i) {
    const double mi = 0.1 * g2[i], vi = 0.001 * static_cast<double>(g2[i]) * g2[i];
    wantp[i] = p2[i] - lr1 * mi / (std::sqrt(vi) + 1e-8);
  }
  ok &= report("adam fused: params", compare(newp, wantp, 1e-6), newp, wantp);
  return ok;
}

// ---------------- perf mode: every conv kernel at the bench shapes, CUDA-event timed ----------------
struct PerfLayer {
  const char* name;
  int cin, cout, k, s;
};
static bool run_perf(int B, int T, int prec, int iters, const char* only = nullptr) {
  const PerfLayer layers[] = {{"striding_conv", 128, 250, 48, 2}, {"inner_conv", 250, 250, 7, 1},
                              {"big_conv_1", 250, 2000, 32, 1},   {"big_conv_2", 2000, 2000, 1, 1},
                              {"output_conv", 2000, 29, 1, 1}};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int t_in = T;
  printf("perf: B=%d T=%d prec=%d (%s)\n", B, T, prec, prec_name(prec));
  for (const auto& L : layers) {
    int T_out, pad_l;
    same_pad(t_in, L.k, L.s, &T_out, &pad_l);
    const int T_alloc = (t_in + L.s - 1) / L.s * L.s;
    const int cip = round64(L.cin), cop = round64(L.cout);
    Dev<uint16_t> xp(static_cast<size_t>(B) * T_alloc * cip * planes_of(prec)), yp(static_cast<size_t>(B) * T_out * cop * planes_of(prec)),
        dxp(static_cast<size_t>(B) * T_alloc * cip * planes_of(prec));
    Dev<uint16_t> wf(static_cast<size_t>(L.k) * cop * cip * planes_of(prec)), wd(static_cast<size_t>(L.k) * cop * cip * planes_of(prec));
    Dev<float> bias(cop), dw(static_cast<size_t>(L.k) * cop * cip), db(cop), probs(static_cast<size_t>(B) * T_out * 32),
        logp(static_cast<size_t>(B) * T_out * 64);
    // fill operands with small random bf16 values (bit patterns via a float pack)
    {
      auto hx = randn(static_cast<size_t>(B) * t_in * L.cin, 1.f);
      Dev<float> dx(hx.size());
      dx.up(hx);
      SLCK(sl_pack_activation(dx.p, xp.p, B, t_in, L.cin, T_alloc, cip, prec, nullptr));
      auto hy = randn(static_cast<size_t>(B) * T_out * L.cout, 1.f);
      Dev<float> dy(hy.size());
      dy.up(hy);
      SLCK(sl_pack_activation(dy.p, yp.p, B, T_out, L.cout, T_out, cop, prec, nullptr));
      auto hw = randn(static_cast<size_t>(L.k) * L.cin * L.cout, 0.05f);
      Dev<float> dwk(hw.size());
      dwk.up(hw);
      SLCK(sl_pack_weights(dwk.p, wf.p, L.k, L.cin, L.cout, cip, cop, prec, nullptr));
    }
    const double flops = 2.0 * L.k * L.cin * L.cout * static_cast<double>(T_out) * B;
    const bool is_out = L.cout == 29;
    const bool skip_timing = only != nullptr && std::string(L.name).find(only) == std::string::npos;
    if (skip_timing) {  // shapes still have to chain
      t_in = T_out;
      continue;
    }
    auto time_it = [&](const char* what, auto fn) -> bool {
      for (int i = 0; i < 2; ++i)
        if (fn() != 0) return false;
      CK(cudaEventRecord(e0));
      for (int i = 0; i < iters; ++i)
        if (fn() != 0) return false;
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= iters;
      printf("  %-14s %-6s %8.3f ms  %7.1f TFLOP/s (algorithmic)\n", L.name, what, ms, flops / ms / 1e9);
      return true;
    };
    bool ok = true;
    Dev<uint16_t> y2(static_cast<size_t>(B) * T_out * cop * planes_of(prec));
    Dev<uint8_t> dgws(L.s == 1 ? sl_conv1d_dgrad_workspace_bytes(B, t_in, L.cin, L.cout, L.k) : 0);
    Dev<uint8_t> pmask(static_cast<size_t>(B) * T_out * cop / 8), pmask_in(static_cast<size_t>(B) * T_alloc * cip / 8);
    CK(cudaMemset(pmask_in.p, 0x5a, pmask_in.n));
    if (is_out)
      ok &= time_it("fwd", [&] {
        return sl_conv1d_fwd(xp.p, wf.p, bias.p, nullptr, nullptr, probs.p, nullptr, logp.p, B, t_in, T_alloc, 0, L.cin, L.cout,
                             L.k, L.s, SL_ACT_SOFTMAX, prec, nullptr);
      });
    else
      ok &= time_it("fwd", [&] {
        return sl_conv1d_fwd(xp.p, wf.p, bias.p, y2.p, pmask.p, nullptr, nullptr, nullptr, B, t_in, T_alloc, 0, L.cin, L.cout, L.k,
                             L.s, SL_ACT_RELU, prec, nullptr);
      });
    if (L.s == 1)
      ok &= time_it("dgrad", [&] {
        return sl_conv1d_dgrad(yp.p, wf.p, pmask_in.p, dxp.p, B, t_in, L.cin, L.cout, L.k, 1, prec, 1.0f,
                               dgws.n > 1 ? dgws.p : nullptr, dgws.n > 1 ? dgws.n : 0, nullptr);
      });
    ok &= time_it("wgrad", [&] {
      return sl_conv1d_wgrad(xp.p, yp.p, dw.p, db.p, B, t_in, T_alloc, L.cin, L.cout, L.k, L.s, prec, 1, 1.0f, nullptr);
    });
    if (!ok) {
      char buf[1024];
      sl_last_error(buf, sizeof buf);
      printf("  perf failure: %s\n", buf);
      return false;
    }
    SLCK(sl_sync_check());
    t_in = T_out;
  }
  return true;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "perf") {
    const int B = argc > 2 ? atoi(argv[2]) : 64;
    const int T = argc > 3 ? atoi(argv[3]) : 1251;
    const int prec = argc > 4 ? atoi(argv[4]) : 1;
    return run_perf(B, T, prec, 5, argc > 5 ? argv[5] : nullptr) ? 0 : 1;
  }
  const std::string filter = argc > 1 ? argv[1] : "";
  auto want = [&](const char* n) { return filter.empty() || std::string(n).find(filter) != std::string::npos; };
  int dev_count = 0;
  CK(cudaGetDeviceCount(&dev_count));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, omp threads %d\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, omp_get_max_threads());
  int failures = 0;
  auto run = [&](const char* name, bool ok) {
    printf("[%s] %s\n", ok ? "PASS" : "FAIL", name);
    fflush(stdout);
    if (!ok) {
      ++failures;
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("sticky CUDA error after %s: %s — aborting\n", name, cudaGetErrorString(e));
        exit(3);
      }
    }
  };
  if (want("pack")) run("pack", test_pack());
  if (want("adam")) run("adam", test_adam());
  if (want("greedy")) run("greedy_golden", test_greedy_golden());

  const ConvCase fwd_cases[] = {
      {"gemm_k1_64x64", 2, 200, 64, 64, 1, 1, SL_ACT_NONE},
      {"gemm_k1_128x128", 1, 128, 128, 128, 1, 1, SL_ACT_NONE},
      {"k3_64x64", 2, 130, 64, 64, 3, 1, SL_ACT_RELU},
      {"inner_k7_250", 2, 300, 250, 250, 7, 1, SL_ACT_RELU},
      {"striding_k48_s2", 2, 301, 128, 250, 48, 2, SL_ACT_RELU},
      {"striding_k48_s2_even", 1, 300, 128, 250, 48, 2, SL_ACT_RELU},
      {"big1_k32_250_2000", 2, 150, 250, 2000, 32, 1, SL_ACT_RELU},
      {"big2_k1_2000_2000", 1, 200, 2000, 2000, 1, 1, SL_ACT_RELU},
      {"out_k1_2000_29", 2, 151, 2000, 29, 1, 1, SL_ACT_SOFTMAX},
      {"out_k1_250_33", 1, 140, 250, 33, 1, 1, SL_ACT_SOFTMAX},
      // > 148 tiles with a remainder: exercises the tail-split (narrow tile) path
      {"tail4_k3_250", 10, 2000, 250, 250, 3, 1, SL_ACT_RELU},
      {"tail2_k1_64_128", 10, 2000, 64, 128, 1, 1, SL_ACT_RELU},
      {"tail_k1_250_2000", 3, 1000, 250, 2000, 1, 1, SL_ACT_RELU},
      // odd number of frame tiles: the last CTA pair has a tile-less CTA
      {"pair_odd_k3_250", 1, 300, 250, 250, 3, 1, SL_ACT_RELU},
      {"pair_odd_k20_250_500", 3, 300, 250, 500, 20, 1, SL_ACT_RELU},
  };
  for (const auto& cc : fwd_cases)
    for (int prec = 1; prec <= 3; ++prec) {
      std::string n = std::string("fwd_") + cc.name + "_p" + std::to_string(prec);
      if (want(n.c_str())) run(n.c_str(), run_conv_fwd(cc, prec));
    }
  const ConvCase dg_cases[] = {
      {"k1_64x64", 2, 200, 64, 64, 1, 1, SL_ACT_NONE},
      {"inner_k7_250", 2, 300, 250, 250, 7, 1, SL_ACT_RELU},
      {"big1_k32_250_2000", 1, 150, 250, 2000, 32, 1, SL_ACT_RELU},
      {"big2_k1_2000_2000", 1, 200, 2000, 2000, 1, 1, SL_ACT_RELU},
      {"out_k1_2000_29", 2, 151, 2000, 29, 1, 1, SL_ACT_RELU},
      {"tail4_k3_250", 10, 2000, 250, 250, 3, 1, SL_ACT_RELU},
      {"tail2_k1_128_64", 10, 2000, 128, 64, 1, 1, SL_ACT_RELU},
      {"splitk_k8_250_2000", 10, 2000, 250, 2000, 8, 1, SL_ACT_RELU},
      {"pair_odd_k3_250", 1, 300, 250, 250, 3, 1, SL_ACT_RELU},
      {"pair_odd_k20_250_500", 3, 300, 250, 500, 20, 1, SL_ACT_RELU},
  };
  for (const auto& cc : dg_cases)
    for (int prec = 1; prec <= 3; ++prec) {
      std::string n = std::string("dgrad_") + cc.name + "_p" + std::to_string(prec);
      if (want(n.c_str())) run(n.c_str(), run_dgrad(cc, prec));
    }
  const ConvCase wg_cases[] = {
      {"k1_64x128", 2, 200, 64, 128, 1, 1, 0},
      {"inner_k7_250", 3, 300, 250, 250, 7, 1, 0},
      {"striding_k48_s2", 2, 301, 128, 250, 48, 2, 0},
      {"big1_k32_250_2000", 1, 150, 250, 2000, 32, 1, 0},
      {"big2_k1_2000_2000", 1, 200, 2000, 2000, 1, 1, 0},
      {"out_k1_2000_29", 2, 151, 2000, 29, 1, 1, 0},
      // 128 input channels: adjacent taps are paired into N = 256 tiles (odd tap counts leave a half-empty pair)
      {"pair_k5_128x250", 2, 300, 128, 250, 5, 1, 0},
      {"pair_k7_s2_128x64", 3, 301, 128, 64, 7, 2, 0},
      {"pair_k2_128x128", 2, 200, 128, 128, 2, 1, 0},
  };
  for (const auto& cc : wg_cases)
    for (int prec = 1; prec <= 3; ++prec) {
      std::string n = std::string("wgrad_") + cc.name + "_p" + std::to_string(prec);
      if (want(n.c_str())) run(n.c_str(), run_wgrad(cc, prec));
    }
  if (want("ctc_small")) run("ctc_small", test_ctc(4, 40, 6, 0, 9, "ctc_small"));
  if (want("ctc_mid")) run("ctc_mid", test_ctc(6, 313, 29, 20, 150, "ctc_mid"));
  if (want("ctc_german")) run("ctc_german", test_ctc(3, 200, 33, 10, 60, "ctc_german"));
  if (want("ctc_long")) run("ctc_long", test_ctc(2, 1500, 29, 500, 700, "ctc_long"));
  // near one-hot distributions (log-probs down to the 1e-8 floor) and (almost) unique alignments: the
  // steepest lattice profiles the linear-domain kernel has to carry through its exponent blocks
  if (want("ctc_peaky")) run("ctc_peaky", test_ctc(4, 300, 29, 40, 140, "ctc_peaky", 12.0f));
  if (want("ctc_tight")) run("ctc_tight", test_ctc(4, 260, 29, 150, 160, "ctc_tight", 12.0f, true));
  if (want("ctc_empty")) run("ctc_empty", test_ctc(3, 50, 29, 0, 1, "ctc_empty"));
  if (want("ctc_bench")) run("ctc_bench", test_ctc(64, 626, 29, 150, 150, "ctc_bench"));
  // long-form shape (BASELINE config 5: 60 s utterances, 16 per GPU at 8 GPUs) and the longest labels built
  if (want("ctc_longform")) run("ctc_longform", test_ctc(16, 3751, 29, 900, 900, "ctc_longform"));
  if (want("ctc_huge")) run("ctc_huge", test_ctc(1, 4200, 29, 1850, 2047, "ctc_huge"));
  printf("selftest finished: %d failure(s)\n", failures);
  return failures ? 1 : 0;
}
