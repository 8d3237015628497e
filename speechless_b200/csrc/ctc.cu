// ctc.cu — CTC loss, gradient and greedy decode.
//
// Replaces K.ctc_batch_cost -> tf.nn.ctc_loss (reference net.py:402-406; blank = V-1,
// grapheme_enconding.py:125-126) and tf.nn.ctc_greedy_decoder (net.py:453-454).
// Semantics restated in SURVEY.md Appendix A.2 / A.3.
//
// Two launches per batch:
//   1. ctc_lattice_halo_kernel — the sequential part.  One CTA (or a cluster of CTAs, for few long
//      utterances) per (utterance, direction): the alpha walk runs t = 0..P-1, the beta walk
//      t = P-1..0 concurrently on other SMs.  The lattice column lives in registers; K time steps run
//      between barriers thanks to a 2K-state halo per warp; log-prob rows are prefetched 32 frames at a
//      time with cp.async; every column is streamed to HBM.  An extra warp of the alpha CTA sorts the
//      label positions by symbol for phase 2.  The loss per utterance is written by the alpha walk.
//   2. ctc_grad_sorted_kernel — the bandwidth part, parallel over every (utterance, frame): one warp
//      per frame folds alpha+beta over the states of each symbol (sorted slots, no atomics) and chains
//      the gradient through log(p+1e-8) and the softmax to the pre-softmax logits, writing the packed
//      bf16 tile the output_conv wgrad/dgrad kernels consume.
// The first-generation kernels (one barrier per step: ctc_alpha_beta_kernel; warp wavefront:
// ctc_alpha_beta_wave_kernel; smem-atomics gradient: ctc_grad_kernel) are kept behind SL_CTC_LEGACY
// for A/B measurements.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace sl {

namespace {

constexpr int CHUNK = 32;  // frames of log-probs staged per cp.async group
constexpr int VP = 64;     // padded symbol row of the logp tensor (V <= 64)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
// L2-only variant: for rows written by other CTAs of the same launch
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2(2^a + 2^b + 2^c), branch free: all -inf -> -inf (lg2(0) = -inf)
__device__ __forceinline__ float lse3_base2(float a, float b, float c) {
  const float m = fmaxf(fmaxf(fmaxf(a, b), c), -1e30f);
  return m + lg2_approx(ex2_approx(a - m) + ex2_approx(b - m) + ex2_approx(c - m));
}

// dir 0: alpha (forward in time); dir 1: beta, computed as the same recurrence on the
// time- and state-reversed problem (the skip condition is symmetric).
// The lattices are kept in BASE-2 logarithms (ex2/lg2 are the native SFU ops; the natural-log
// inputs are rescaled once per frame), SPT consecutive states per thread.
template <int SPT>
__global__ void ctc_alpha_beta_kernel(const float* __restrict__ logp,
                                      const int32_t* __restrict__ labels,
                                      const int32_t* __restrict__ input_len,
                                      const int32_t* __restrict__ label_len,
                                      float* __restrict__ loss, float* __restrict__ beta_loss,
                                      float* __restrict__ alpha, float* __restrict__ beta, int T,
                                      int L_max, int blank, int S_stride, int col_stride) {
  extern __shared__ uint8_t smem_raw[];
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();  // logp comes from the output_conv kernel right before
  const int dir = blockIdx.x & 1;
  const int b = blockIdx.x >> 1;
  const int tid = threadIdx.x;
  const int L = label_len[b];
  const int P = min(input_len[b], T);
  const int S = 2 * L + 1;

  // smem carve-up
  float* lp_s = reinterpret_cast<float*>(smem_raw);              // ring of 2*CHUNK rows x VP
  float* col = lp_s + 2 * CHUNK * VP;                            // [2][col_stride], 2 leading pads
  int* ext = reinterpret_cast<int*>(col + 2 * col_stride);       // [S_stride]

  // extended label sequence in walking order (reversed for beta)
  for (int s = tid; s < S_stride; s += blockDim.x) {
    int e = blank;
    if (s < S) {
      const int so = dir ? (S - 1 - s) : s;
      if (so & 1) e = labels[static_cast<size_t>(b) * L_max + (so >> 1)];
    }
    ext[s] = e;
  }
  for (int i = tid; i < 2 * col_stride; i += blockDim.x) col[i] = -INFINITY;

  const float* lp_b = logp + static_cast<size_t>(b) * T * VP;
  auto issue_chunk = [&](int c) {
    // rows tt = c*CHUNK .. +CHUNK-1 in walking order -> ring rows (c & 1) * CHUNK ..
    float* dst = lp_s + (c & 1) * CHUNK * VP;
    for (int i = tid; i < CHUNK * (VP / 4); i += blockDim.x) {
      const int r = i / (VP / 4), piece = i % (VP / 4);
      const int tt = c * CHUNK + r;
      if (tt < P) {
        const int t = dir ? (P - 1 - tt) : tt;
        cp_async16(dst + r * VP + piece * 4, lp_b + static_cast<size_t>(t) * VP + piece * 4);
      }
    }
    cp_async_commit();
  };
  issue_chunk(0);
  issue_chunk(1);
  __syncthreads();

  // per-thread state bookkeeping: symbol, skip permission, smem / global offsets
  const int s0 = tid * SPT;
  int my_e[SPT];
  bool my_skip[SPT], my_on[SPT];
  int my_out[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int s = s0 + i;
    my_on[i] = s < S;
    my_e[i] = s < S_stride ? ext[s] : blank;
    my_skip[i] = (s >= 2 && s < S) ? (ext[s] != blank && ext[s] != ext[s - 2]) : false;
    my_out[i] = dir ? (S - 1 - s) : s;
  }

  float* out = (dir ? beta : alpha) + static_cast<size_t>(b) * T * S_stride +
               (dir ? static_cast<size_t>(P > 0 ? P - 1 : 0) * S_stride : 0);
  const ptrdiff_t out_step = dir ? -static_cast<ptrdiff_t>(S_stride) : static_cast<ptrdiff_t>(S_stride);
  float* prev = col + col_stride + 2;  // column of step tt-1 (initially all -inf)
  float* cur = col + 2;

  for (int tt = 0; tt < P; ++tt) {
    if ((tt & (CHUNK - 1)) == 0) {
      cp_async_wait<1>();
      __syncthreads();
    }
    const float* lp_row = lp_s + (tt & (2 * CHUNK - 1)) * VP;
    float val[SPT];
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const int s = s0 + i;
        const float em = lp_row[my_e[i]] * LOG2E;
        const float a0 = prev[s];
        const float a1 = prev[s - 1];
        const float a2 = my_skip[i] ? prev[s - 2] : -INFINITY;
        val[i] = lse3_base2(a0, a1, a2) + em;
        if (tt == 0) val[i] = (s <= 1) ? em : -INFINITY;
      }
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        if (my_on[i]) {
          cur[s0 + i] = val[i];
          out[my_out[i]] = val[i];
        }
      }
    out += out_step;
    __syncthreads();
    if ((tt & (CHUNK - 1)) == CHUNK - 1) issue_chunk(tt / CHUNK + 2);
    float* tmp = prev;
    prev = cur;
    cur = tmp;
  }
  cp_async_wait<0>();

  if (tid == 0) {
    float l = INFINITY;
    if (P > 0) {
      const float a = prev[S - 1];  // prev = column of the last processed step
      const float c = S >= 2 ? prev[S - 2] : -INFINITY;
      const float m = fmaxf(a, c);
      l = (m == -INFINITY) ? INFINITY : -(m + log2f(exp2f(a - m) + exp2f(c - m))) * LN2;
    }
    if (dir == 0)
      loss[b] = l;
    else
      beta_loss[b] = l;
  }
}

// ---------------------------------------------------------------------------------------------
// Halo-blocked variant (the default): K time steps between block barriers.
//
// The lattice column lives in REGISTERS: lane l of warp w holds the SPT consecutive states
// base_w + l*SPT ..., where base_w = w*OWN - 2K and OWN = 32*SPT - 2K.  The first 2K slots of a
// warp are a HALO that overlaps the last 2K owned states of warp w-1.  A state at step t depends
// on itself and its two left neighbours at step t-1, so after k steps without any exchange the
// slots >= 2k of a warp are still exact: K steps run with one warp shuffle per step and no
// shared-memory round trip or barrier; then every warp publishes its owned states to a
// double-buffered smem column, ONE barrier, and the halos are reloaded.  The redundant halo work
// (2K of 32*SPT slots) buys a K-fold cut of the barrier + smem latency of the one-barrier-per-step
// kernel above.  ncu on the first version of this kernel showed what bounds such a recurrence with
// 1-2 warps per scheduler: not the SFU, but the NUMBER of dependent instructions per step (~2 cycles
// of fixed-latency "wait" stall per instruction).  Hence the step is written for a minimal count:
//  * "log 0" is the finite sentinel NEG = -1e30 (absorbing under +emission, 2^(NEG - m) = 0), so no
//    NaN guard is needed for (-inf) - (-inf);
//  * "state does not exist" and "skip transition not allowed" are additive biases (0 or NEG) folded
//    into the emission / the s-2 term — no predicates, no selects;
//  * SPT is even and base_w is even, so slot parity = state parity in walking order (S is odd, so
//    this also holds for the reversed beta problem): even slots are blank states (two-term
//    log-sum-exp: 1 ex2 + 1 lg2), odd slots label states (three terms: 2 ex2 + 1 lg2 — the largest
//    term is 2^0 and is not exponentiated);
//  * emission rows are read with immediate offsets from a per-block base pointer; full K-blocks run
//    without per-step bounds checks; the direction is a template parameter of the walk.
// An extra warp of the alpha CTA counting-sorts the label positions by symbol for the gradient
// kernel (see ctc_grad_sorted_kernel) while the compute warps walk the lattice.
constexpr float NEG = -1e30f;
constexpr int FLAG_STRIDE = 32;  // ints: every progress flag has a 128-byte line of its own
constexpr int ROLE_COUNTERS = 96 + 256;  // ints behind the flags: role counters of the fused launch (+ one per SM)

// progress flags between the lattice walkers and the gradient CTAs of one launch
__device__ __forceinline__ void st_release_cta_shared(int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(ptx::smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr int SORT_EXTRA = 96;  // ints after the packed labels: seg[VP + 1] (+ padding)

// Stable counting sort of the label positions of one utterance by symbol, by ONE warp.
// packed[i] = symbol | slot << 8 (slot = position of label i in the symbol-sorted order),
// seg[v] = first slot of symbol v, seg[VP] = L.  cnt: VP ints of scratch smem.
__device__ __forceinline__ void sort_labels_by_symbol(const int32_t* __restrict__ lab, int L, int* __restrict__ packed,
                                                      int* __restrict__ seg, int* cnt, int lane) {
  cnt[lane] = 0;
  cnt[lane + 32] = 0;
  __syncwarp();
  for (int i0 = 0; i0 < L; i0 += 32) {
    const int i = i0 + lane;
    // (distinct dummies past the end; symbols are clamped into the table: the host rejects labels outside
    // [0, V - 2] before the launch, this only keeps a bad direct caller inside shared memory)
    const int v = i < L ? min(max(lab[i], 0), VP - 1) : -1 - lane;
    const unsigned same = __match_any_sync(0xffffffffu, v);
    if (i < L) {
      const int before = cnt[v];
      packed[i] = v | ((before + __popc(same & ((1u << lane) - 1))) << 8);
      __syncwarp(same);
      if ((same >> lane) == 1u) cnt[v] = before + __popc(same);  // highest lane of the group
    }
    __syncwarp();
  }
  // exclusive prefix over the symbols (V <= 64: two per lane)
  const int c0 = cnt[lane], c1 = cnt[lane + 32];
  int inc0 = c0, inc1 = c1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u0 = __shfl_up_sync(0xffffffffu, inc0, o), u1 = __shfl_up_sync(0xffffffffu, inc1, o);
    if (lane >= o) {
      inc0 += u0;
      inc1 += u1;
    }
  }
  const int total0 = __shfl_sync(0xffffffffu, inc0, 31);
  __syncwarp();
  cnt[lane] = inc0 - c0;
  cnt[lane + 32] = total0 + inc1 - c1;
  seg[lane] = inc0 - c0;
  seg[lane + 32] = total0 + inc1 - c1;
  if (lane == 31) seg[VP] = total0 + inc1;
  __syncwarp();
  for (int i = lane; i < L; i += 32) {  // (each lane revisits the entries it wrote itself)
    const int pk = packed[i];
    packed[i] = (pk & 255) | (((pk >> 8) + cnt[pk & 255]) << 8);
  }
}

// push_rank >= 0 (cluster mode, last compute warp of a CTA): the lanes holding the warp's last 2K
// owned states also store them into the halo slots of the next CTA of the cluster (DSMEM).
template <int SPT, int K, int DIR, bool CL>
__device__ __forceinline__ void halo_walk(const float* lp_s, float* col, int col_stride, const float* lp_b, int P,
                                          int ls0, int lane, int tid, int nthreads, const float (&skipb)[SPT],
                                          const float (&onb)[SPT], const float* const (&em_ptr)[SPT], float* out,
                                          ptrdiff_t out_step, unsigned st_mask, int push_rank, int own_per_cta,
                                          int* smem_progress) {
  constexpr int HALO = 2 * K;
  const bool owner = lane * SPT >= HALO;
  const bool pusher = CL && push_rank >= 0 && lane * SPT >= 32 * SPT - HALO;
  auto barrier = [&]() {
    if constexpr (CL)
      ptx::cluster_sync_all();
    else
      bar_sync_named(1, nthreads);
  };
  auto issue_chunk = [&](int c) {
    float* dst = const_cast<float*>(lp_s) + (c & 1) * CHUNK * VP;
    for (int i = tid; i < CHUNK * (VP / 4); i += nthreads) {
      const int r = i / (VP / 4), piece = i % (VP / 4);
      const int tt = c * CHUNK + r;
      if (tt < P) {
        const int t = DIR ? (P - 1 - tt) : tt;
        cp_async16(dst + r * VP + piece * 4, lp_b + static_cast<size_t>(t) * VP + piece * 4);
      }
    }
    cp_async_commit();
  };
  issue_chunk(0);
  issue_chunk(1);

  float prev[SPT];
  const float* ep[SPT];  // emission pointers of the current K-block's first row
  // one step: prev -> prev; k = row within the block (an immediate offset in the unrolled full blocks)
  auto step = [&](int k) {
    const float left = __shfl_up_sync(0xffffffffu, prev[SPT - 1], 1);  // lane 0: own value, stays inside the halo
    float cur[SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const float emb = fmaf(ep[i][k * VP], LOG2E, onb[i]);
      const float a0 = prev[i];
      const float a1 = i >= 1 ? prev[i - 1] : left;
      float m, sum;
      if (i & 1) {
        const float a2 = (i >= 2 ? prev[i - 2] : left) + skipb[i];
        const float hi = fmaxf(a0, a1), lo = fminf(a0, a1);
        m = fmaxf(hi, a2);
        const float md = fminf(hi, a2);  // {a0,a1,a2} \ {max} = {lo, md}
        sum = (1.0f + ex2_approx(lo - m)) + ex2_approx(md - m);
      } else {
        m = fmaxf(a0, a1);
        sum = 1.0f + ex2_approx(fminf(a0, a1) - m);
      }
      cur[i] = (m + emb) + lg2_approx(sum);
    }
    if (DIR == 0) {
      if (st_mask) {
        if constexpr (SPT == 2) {
          *reinterpret_cast<float2*>(out) = make_float2(cur[0], cur[1]);
        } else {
#pragma unroll
          for (int i = 0; i < SPT; i += 4)
            *reinterpret_cast<float4*>(out + i) = make_float4(cur[i], cur[i + 1], cur[i + 2], cur[i + 3]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < SPT; ++i)
        if ((st_mask >> i) & 1u) out[-i] = cur[i];
    }
    out += out_step;
#pragma unroll
    for (int i = 0; i < SPT; ++i) prev[i] = cur[i];
  };

  for (int t0 = 0; t0 < P; t0 += K) {
    const int kb = t0 / K;
    if ((t0 & (CHUNK - 1)) == 0) cp_async_wait<0>();  // the chunk issued one chunk ago has long landed
    barrier();  // owned states of block kb-1 published; chunk visible; ring slot free
    // every lattice row of the steps < t0 has been stored by its thread: tell the publisher warp
    if (!CL && smem_progress != nullptr && tid == 0 && t0 > 0) st_release_cta_shared(smem_progress, t0);
    if ((t0 & (CHUNK - 1)) == 0 && t0 > 0) issue_chunk(t0 / CHUNK + 1);
    const float* pc = col + ((kb + 1) & 1) * col_stride + HALO + ls0;
#pragma unroll
    for (int i = 0; i < SPT; ++i) prev[i] = pc[i];
#pragma unroll
    for (int i = 0; i < SPT; ++i) ep[i] = em_ptr[i] + (t0 & (2 * CHUNK - 1)) * VP;  // (a K-block never wraps the ring)
    if (t0 + K <= P) {
#pragma unroll
      for (int k = 0; k < K; ++k) step(k);
    } else {
      for (int k = 0; t0 + k < P; ++k) step(k);
    }
    if (owner) {
      float* wc = col + (kb & 1) * col_stride + HALO + ls0;
#pragma unroll
      for (int i = 0; i < SPT; ++i) wc[i] = prev[i];
      if (pusher) {
        // same buffer, local index of the next CTA: its state base is own_per_cta further on
        const uint32_t remote = ptx::mapa_u32(ptx::smem_u32(wc - own_per_cta), static_cast<uint32_t>(push_rank));
#pragma unroll
        for (int i = 0; i < SPT; ++i)
          asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote + 4u * i), "f"(prev[i]) : "memory");
      }
    }
  }
  cp_async_wait<0>();
  barrier();
  if (!CL && smem_progress != nullptr && tid == 0) st_release_cta_shared(smem_progress, P);
}

// One (utterance, direction) of the lattice phase, run by a whole CTA (or by a cluster of CTAs, CL).
// progress != nullptr (fused loss + gradient launch): the spare warp also PUBLISHES how many time steps of
// this walk are complete in global memory — progress[2 b + dir] — for the gradient CTAs of the same launch.
// The walking warps never wait for the publication: walker thread 0 drops the step count into shared memory
// after each K-block barrier (st.release.cta), the publisher warp picks it up (ld.acquire.cta), issues the
// device-scope fence that makes the rows stored before that barrier visible, and stores the flag.
template <int SPT, int K, bool CL>
__device__ __forceinline__ void lattice_cta_body(uint8_t* smem_raw, int unit, const float* __restrict__ logp,
                                                 const int32_t* __restrict__ labels,
                                                 const int32_t* __restrict__ input_len,
                                                 const int32_t* __restrict__ label_len, float* __restrict__ loss,
                                                 float* __restrict__ beta_loss, float* __restrict__ alpha,
                                                 float* __restrict__ beta, int* __restrict__ sort_ws, int T,
                                                 int L_max, int blank, int S_stride, int col_stride,
                                                 int* __restrict__ progress, int dbg = 0);

template <int SPT, int K, bool CL>
__global__ void __launch_bounds__(1024)
    ctc_lattice_halo_kernel(const float* __restrict__ logp, const int32_t* __restrict__ labels,
                            const int32_t* __restrict__ input_len, const int32_t* __restrict__ label_len,
                            float* __restrict__ loss, float* __restrict__ beta_loss,
                            float* __restrict__ alpha, float* __restrict__ beta, int* __restrict__ sort_ws,
                            int T, int L_max, int blank, int S_stride, int col_stride) {
  extern __shared__ uint8_t smem_raw[];
  ptx::pdl_launch_dependents();
  const int csize = CL ? static_cast<int>(ptx::cluster_nctarank()) : 1;
  lattice_cta_body<SPT, K, CL>(smem_raw, static_cast<int>(blockIdx.x) / csize, logp, labels, input_len, label_len, loss,
                               beta_loss, alpha, beta, sort_ws, T, L_max, blank, S_stride, col_stride, nullptr);
}

template <int SPT, int K, bool CL>
__device__ __forceinline__ void lattice_cta_body(uint8_t* smem_raw, int unit, const float* __restrict__ logp,
                                                 const int32_t* __restrict__ labels,
                                                 const int32_t* __restrict__ input_len,
                                                 const int32_t* __restrict__ label_len, float* __restrict__ loss,
                                                 float* __restrict__ beta_loss, float* __restrict__ alpha,
                                                 float* __restrict__ beta, int* __restrict__ sort_ws, int T,
                                                 int L_max, int blank, int S_stride, int col_stride,
                                                 int* __restrict__ progress, int dbg) {
  static_assert(SPT % 2 == 0 && CHUNK % K == 0 && 2 * K < 32 * SPT, "slot parity / chunk alignment");
  constexpr int HALO = 2 * K;
  constexpr int OWN = 32 * SPT - HALO;
  // cluster mode: the CTAs of a cluster split the warps (= the state range) of one (utterance, direction)
  const int crank = CL ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int csize = CL ? static_cast<int>(ptx::cluster_nctarank()) : 1;
  const int dir = unit & 1;
  const int b = unit >> 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthreads = blockDim.x - 32;  // compute threads; the last warp is the label sorter
  const int own_per_cta = (nthreads >> 5) * OWN;
  const int L = label_len[b];
  const int P = min(input_len[b], T);
  const int S = 2 * L + 1;

  float* lp_s = reinterpret_cast<float*>(smem_raw);  // ring of 2*CHUNK rows x VP
  float* col = lp_s + 2 * CHUNK * VP;                // [2][col_stride], index = state - crank*own_per_cta + HALO
  int* sort_scratch = reinterpret_cast<int*>(col + 2 * col_stride);  // [VP]
  int* smem_progress = progress != nullptr ? sort_scratch + VP : nullptr;  // steps of this walk stored so far
  if (progress != nullptr) {
    if (tid == 0) *smem_progress = 0;
    __syncthreads();
  }

  if (tid >= nthreads) {
    // label sorter (does not take part in the walk's CTA barriers); the labels are inputs of the
    // step, not products of the kernel before, so it does not wait for it either
    if (dir == 0 && crank == 0) {
      const int L_pad = (L_max + 31) & ~31;
      int* packed = sort_ws + static_cast<size_t>(b) * (L_pad + SORT_EXTRA);
      sort_labels_by_symbol(labels + static_cast<size_t>(b) * L_max, L, packed, packed + L_pad, sort_scratch, lane);
    }
    // ... and publisher of the walk's progress (the sorted labels above are covered by the first fence)
    auto publish = [&](int done) {
      __threadfence();
      if (lane == 0) st_relaxed_gpu(progress + unit * FLAG_STRIDE, done);
    };
    if constexpr (CL) {  // cluster barriers count every thread: one per K-block and the final one
      for (int t0 = 0; t0 < P; t0 += K) {
        ptx::cluster_sync_all();  // every CTA of the cluster has stored the rows of the steps < t0
        if (progress != nullptr && crank == 0 && t0 > 0) publish(t0);
      }
      ptx::cluster_sync_all();
      if (progress != nullptr && crank == 0) publish(P);
    } else if (progress != nullptr) {
      int published = 0;
      unsigned spins = 0;
      while (published < P) {
        const int done = ld_acquire_cta_shared(smem_progress);
        if (done > published) {
          publish(done);
          published = done;
          spins = 0;
        } else {
          __nanosleep(100);
          if (++spins > (1u << 24)) __trap();  // a protocol bug must fail, not hang
        }
      }
      if (P <= 0) publish(0);
    }
    return;
  }

  // virtual column of step -1: state 0 holds log 1, so the generic recurrence yields
  // alpha_0(0) = em(0), alpha_0(1) = em(1), everything else "log 0"
  for (int i = tid; i < 2 * col_stride; i += nthreads)
    col[i] = (i == col_stride + HALO && crank == 0) ? 0.0f : NEG;

  const int ls0 = warp * OWN - HALO + lane * SPT;  // state index relative to this CTA's first owned state
  const int s0 = crank * own_per_cta + ls0;        // even; negative inside the very first halo
  float skipb[SPT], onb[SPT];
  const float* em_ptr[SPT];
  unsigned st_mask = 0;
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int s = s0 + i;
    const bool on = s >= 0 && s < S;
    auto symbol = [&](int st) {  // extended label sequence in walking order (reversed for beta)
      const int so = dir ? (S - 1 - st) : st;
      return (so & 1) ? labels[static_cast<size_t>(b) * L_max + (so >> 1)] : blank;
    };
    const int e = (on && (i & 1)) ? symbol(s) : blank;
    const bool skip = (on && (i & 1) && s >= 3) ? (e != symbol(s - 2)) : false;
    skipb[i] = skip ? 0.0f : NEG;
    onb[i] = on ? 0.0f : NEG;
    em_ptr[i] = lp_s + e;
    if (on && lane * SPT >= HALO) st_mask |= 1u << i;
  }
  ptx::pdl_wait();  // logp comes from the output_conv kernel right before

  const float* lp_b = logp + static_cast<size_t>(b) * T * VP;
  float* lat = (dir ? beta : alpha) + static_cast<size_t>(b) * T * S_stride;
  const int push_rank = (CL && crank + 1 < csize && warp == (nthreads >> 5) - 1) ? crank + 1 : -1;
  if (dir == 0)
    halo_walk<SPT, K, 0, CL>(lp_s, col, col_stride, lp_b, P, ls0, lane, tid, nthreads, skipb, onb, em_ptr, lat + s0,
                             static_cast<ptrdiff_t>(S_stride), st_mask, push_rank, own_per_cta, smem_progress);
  else
    halo_walk<SPT, K, 1, CL>(lp_s, col, col_stride, lp_b, P, ls0, lane, tid, nthreads, skipb, onb, em_ptr,
                             lat + static_cast<size_t>(P > 0 ? P - 1 : 0) * S_stride + (S - 1 - s0),
                             -static_cast<ptrdiff_t>(S_stride), st_mask, push_rank, own_per_cta, smem_progress);

  // the CTA owning the last state reports the loss (state S-2 is its own or sits in its halo)
  if (tid == 0 && (S - 1) / own_per_cta == crank) {
    float l = INFINITY;
    if (P > 0) {
      const float* fc = col + (((P + K - 1) / K - 1) & 1) * col_stride + HALO - crank * own_per_cta;
      const float a = fc[S - 1];
      const float c = S >= 2 ? fc[S - 2] : NEG;
      const float m = fmaxf(a, c);
      l = (m < -1e29f) ? INFINITY : -(m + log2f(exp2f(a - m) + exp2f(c - m))) * LN2;
    }
    if (dir == 0)
      loss[b] = l;
    else
      beta_loss[b] = l;
  }
}

// ---------------------------------------------------------------------------------------------
// Wavefront variant of the alpha/beta recurrence: no block-wide barrier inside the time loop.
// (Opt-in experiment, SL_CTC_WAVE=1: correct, but slower than the barrier kernel — see the launcher.)
//
// Warp w owns the 32*SPT consecutive lattice states [w*32*SPT, ...), lane l the SPT states
// s0 = (w*32 + l)*SPT ...  A state needs the previous column's values of itself and of its two left
// neighbours: the lane's own registers plus the last two states of lane l-1 (warp shuffle); lane 0
// takes them from the previous warp through a small smem ring (edge values + progress counters), so
// the warps form a pipeline skewed by one time step and nobody waits in steady state.  A loader
// warp (the last one) streams the log-prob rows into a 128-frame smem ring with cp.async and
// publishes how many rows are ready; it reuses a ring slot only after every compute warp has passed
// it.  Spins are bounded (trap, never hang).
constexpr int LP_RING = 128;  // frames of log-probs resident in smem
constexpr int EDGE_RING = 16;

__device__ __forceinline__ int ld_volatile_s32(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_s32(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(ptx::smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void spin_until_at_least(const int* p, int target) {
  unsigned spins = 0;
  while (ld_volatile_s32(p) < target)
    if (++spins > (1u << 26)) __trap();  // a protocol bug must fail, not hang
}

template <int SPT>
__global__ void ctc_alpha_beta_wave_kernel(const float* __restrict__ logp,
                                           const int32_t* __restrict__ labels,
                                           const int32_t* __restrict__ input_len,
                                           const int32_t* __restrict__ label_len,
                                           float* __restrict__ loss, float* __restrict__ beta_loss,
                                           float* __restrict__ alpha, float* __restrict__ beta, int T,
                                           int L_max, int blank, int S_stride) {
  extern __shared__ uint8_t smem_raw[];
  const int dir = blockIdx.x & 1;
  const int b = blockIdx.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_compute = (blockDim.x >> 5) - 1;  // the last warp is the loader
  const int L = label_len[b];
  const int P = min(input_len[b], T);
  const int S = 2 * L + 1;
  const int n_active = (S + 32 * SPT - 1) / (32 * SPT);  // warps that own at least one state

  float* lp_s = reinterpret_cast<float*>(smem_raw);                         // [LP_RING][VP]
  // edge ring: {value of state s_last-1, value of s_last, step + 1, 0} written with ONE 16-byte
  // store, so the payload and its sequence tag become visible together and the hot loop needs no
  // memory fence (a fence would also wait for the global lattice stores in flight)
  uint4* edge = reinterpret_cast<uint4*>(lp_s + LP_RING * VP);              // [n_compute][EDGE_RING]
  float* final_col = reinterpret_cast<float*>(edge + n_compute * EDGE_RING);  // [n_compute * 32 * SPT]
  int* done = reinterpret_cast<int*>(final_col + n_compute * 32 * SPT);     // [n_compute] steps completed
  int* lp_ready = done + n_compute;                                         // rows of logp available

  for (int i = threadIdx.x; i <= n_compute; i += blockDim.x) done[i] = 0;  // (includes lp_ready)
  for (int i = threadIdx.x; i < n_compute * EDGE_RING; i += blockDim.x) edge[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  const float* lp_b = logp + static_cast<size_t>(b) * T * VP;
  if (warp == n_compute) {
    // ===================== loader warp =====================
    const int chunks = (P + CHUNK - 1) / CHUNK;
    for (int c = 0; c < chunks; ++c) {
      if (c >= LP_RING / CHUNK) {
        // the ring slot of chunk c still holds chunk c - 4: every active warp must be past it
        const int need = (c - LP_RING / CHUNK + 1) * CHUNK;
        for (int w = lane; w < n_active; w += 32) spin_until_at_least(&done[w], need);
        __syncwarp();
      }
      float* dst = lp_s + (c % (LP_RING / CHUNK)) * CHUNK * VP;
      for (int i = lane; i < CHUNK * (VP / 4); i += 32) {
        const int r = i / (VP / 4), piece = i % (VP / 4);
        const int tt = c * CHUNK + r;
        if (tt < P) {
          const int t = dir ? (P - 1 - tt) : tt;
          cp_async16(dst + r * VP + piece * 4, lp_b + static_cast<size_t>(t) * VP + piece * 4);
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      __threadfence_block();
      __syncwarp();
      if (lane == 0) st_volatile_s32(lp_ready, min(P, (c + 1) * CHUNK));
    }
  } else if (warp < n_active) {
    // ===================== compute warps =====================
    const int s0 = (warp * 32 + lane) * SPT;
    int my_e[SPT], my_out[SPT];
    bool my_skip[SPT], my_on[SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const int s = s0 + i;
      my_on[i] = s < S;
      auto symbol = [&](int st) {  // extended label sequence in walking order (reversed for beta)
        const int so = dir ? (S - 1 - st) : st;
        return (so & 1) ? labels[static_cast<size_t>(b) * L_max + (so >> 1)] : blank;
      };
      my_e[i] = my_on[i] ? symbol(s) : blank;
      my_skip[i] = (my_on[i] && s >= 2) ? (my_e[i] != blank && my_e[i] != symbol(s - 2)) : false;
      my_out[i] = dir ? (S - 1 - s) : s;
    }
    float prev[SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) prev[i] = -INFINITY;
    float* out = (dir ? beta : alpha) + static_cast<size_t>(b) * T * S_stride +
                 (dir ? static_cast<size_t>(P > 0 ? P - 1 : 0) * S_stride : 0);
    const ptrdiff_t out_step = dir ? -static_cast<ptrdiff_t>(S_stride) : static_cast<ptrdiff_t>(S_stride);
    const bool last_warp = warp == n_active - 1;

    for (int tt = 0; tt < P; ++tt) {
      if ((tt & (CHUNK - 1)) == 0) spin_until_at_least(lp_ready, min(P, tt + CHUNK));
      const float* lp_row = lp_s + (tt & (LP_RING - 1)) * VP;
      float em[SPT];
#pragma unroll
      for (int i = 0; i < SPT; ++i) em[i] = lp_row[my_e[i]] * LOG2E;
      // values of states s0-1, s0-2 at the previous step.  Every poll below is executed by the
      // whole warp on one address (warp-uniform trip counts): single-lane spin loops leave the
      // warp diverged, and every later shuffle then pays for a re-convergence.
      float left1 = __shfl_up_sync(0xffffffffu, prev[SPT - 1], 1);
      float left0 = __shfl_up_sync(0xffffffffu, prev[SPT - 2 >= 0 ? SPT - 2 : 0], 1);
      float edge0 = -INFINITY, edge1 = -INFINITY;
      if (warp > 0 && tt > 0) {
        // poll the slot of step tt-1 until its tag says so (payload and tag share one 16-byte word)
        const uint32_t slot = ptx::smem_u32(&edge[(warp - 1) * EDGE_RING + ((tt - 1) & (EDGE_RING - 1))]);
        uint4 e;
        unsigned spins = 0;
        do {
          asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                       : "r"(slot)
                       : "memory");
          if (++spins > (1u << 26)) __trap();
        } while (e.z != static_cast<uint32_t>(tt));
        edge0 = __uint_as_float(e.x);
        edge1 = __uint_as_float(e.y);
      }
      left0 = lane == 0 ? edge0 : left0;
      left1 = lane == 0 ? edge1 : left1;
      float cur[SPT];
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const float a0 = prev[i];
        const float a1 = i >= 1 ? prev[i - 1] : left1;
        const float a2raw = i >= 2 ? prev[i - 2] : (i == 1 ? left1 : left0);
        const float a2 = my_skip[i] ? a2raw : -INFINITY;
        cur[i] = lse3_base2(a0, a1, a2) + em[i];
        if (tt == 0) cur[i] = (s0 + i <= 1) ? em[i] : -INFINITY;
        if (!my_on[i]) cur[i] = -INFINITY;
      }
      // publish this warp's last two states for the next warp, then report the step as done (the
      // report lets the previous warp overwrite the slot lane 0 just read, hence the __syncwarp)
      if (!last_warp) {
        // slot tt % R was read by the next warp during its step tt - R + 1 (uniform poll)
        if (tt >= EDGE_RING) spin_until_at_least(&done[warp + 1], tt - EDGE_RING + 2);
        if (lane == 31) {
          const uint32_t slot = ptx::smem_u32(&edge[warp * EDGE_RING + (tt & (EDGE_RING - 1))]);
          asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot),
                       "r"(__float_as_uint(cur[SPT - 2 >= 0 ? SPT - 2 : 0])), "r"(__float_as_uint(cur[SPT - 1])),
                       "r"(static_cast<uint32_t>(tt + 1)), "r"(0u)
                       : "memory");
        }
      }
      __syncwarp();
      if (lane == 31) st_volatile_s32(&done[warp], tt + 1);  // progress: back-pressure and ring reuse only
#pragma unroll
      for (int i = 0; i < SPT; ++i)
        if (my_on[i]) out[my_out[i]] = cur[i];
      out += out_step;
#pragma unroll
      for (int i = 0; i < SPT; ++i) prev[i] = cur[i];
    }
#pragma unroll
    for (int i = 0; i < SPT; ++i) final_col[s0 + i] = prev[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = INFINITY;
    if (P > 0) {
      const float a = final_col[S - 1];
      const float c = S >= 2 ? final_col[S - 2] : -INFINITY;
      const float m = fmaxf(a, c);
      l = (m == -INFINITY) ? INFINITY : -(m + log2f(exp2f(a - m) + exp2f(c - m))) * LN2;
    }
    if (dir == 0)
      loss[b] = l;
    else
      beta_loss[b] = l;
  }
}

// One warp per (utterance, frame).  grad wrt logits z of  grad_scale * sum_b loss_b:
//   u = log(p+eps), lp = log_softmax(u);  g_u(v) = exp(lp_v) - occ(v)
//   occ(v) = sum_{s: e[s]=v} exp(alpha_t(s) + beta_t(s) - lp_t(v) + loss_b)
//   dL/dp_v = g_u(v)/(p_v+eps);  dL/dz_v = p_v (dL/dp_v - sum_w p_w dL/dp_w)
// alpha/beta arrive as base-2 logs; each lane reads 4 consecutive states (float4) per pass.
__global__ void ctc_grad_kernel(const float* __restrict__ logp, const float* __restrict__ probs,
                                const int32_t* __restrict__ labels,
                                const int32_t* __restrict__ input_len,
                                const int32_t* __restrict__ label_len,
                                const float* __restrict__ loss, const float* __restrict__ alpha,
                                const float* __restrict__ beta, __nv_bfloat16* __restrict__ dz_packed,
                                float* __restrict__ dz_f32, float grad_scale, int T, int V,
                                int L_max, int blank, int S_stride, int planes, int fp16,
                                int frames_per_block) {
  extern __shared__ uint8_t smem_raw[];
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();  // alpha / beta / loss come from the lattice kernel right before
  int* lab_s = reinterpret_cast<int*>(smem_raw);                    // [L_max]
  float* bins = reinterpret_cast<float*>(lab_s + ((L_max + 31) & ~31));  // [warps][VP]
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int L = label_len[b];
  const int P = min(input_len[b], T);
  const int S = 2 * L + 1;
  for (int i = threadIdx.x; i < L; i += blockDim.x) lab_s[i] = labels[static_cast<size_t>(b) * L_max + i];
  __syncthreads();
  const float loss_b = loss[b];
  const float loss2 = loss_b * LOG2E;
  float* my_bins = bins + warp * (2 * VP);  // [VP] occupancy bins | [VP] log-prob row
  const int t_begin = blockIdx.x * frames_per_block;
  const int t_end = min(t_begin + frames_per_block, T);
  const int row_elems = planes * 64;

  for (int t = t_begin + warp; t < t_end; t += nwarps) {
    const size_t ro = static_cast<size_t>(b) * T + t;
    float dz[2] = {0.f, 0.f};  // symbols lane and lane + 32
    if (t < P && isfinite(loss_b)) {
      // issue every independent load of this frame up front: the symbol row (to smem), the
      // probabilities, then the lattice rows
      const float* lp_g = logp + ro * VP;
      const float lp_mine[2] = {lp_g[lane], lp_g[lane + 32]};
      float pv_mine[2] = {0.f, 0.f};
      if (lane < V) pv_mine[0] = probs[ro * V + lane];
      if (lane + 32 < V) pv_mine[1] = probs[ro * V + lane + 32];
      float* lp_row = my_bins + VP;  // per-warp copy of the log-prob row (base 2)
      my_bins[lane] = 0.f;
      my_bins[lane + 32] = 0.f;
      lp_row[lane] = lp_mine[0] * LOG2E;
      lp_row[lane + 32] = lp_mine[1] * LOG2E;
      __syncwarp();
      const float4* a_row = reinterpret_cast<const float4*>(alpha + ro * S_stride);
      const float4* b_row = reinterpret_cast<const float4*>(beta + ro * S_stride);
      const float lp_blank2 = lp_row[blank];
      float blank_acc = 0.f;
      for (int base = 0; base < S; base += 512) {
        // up to 4 passes of 128 states issued together (8 x 16-byte loads in flight per lane)
        float4 av[4], bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int s = base + j * 128 + lane * 4;
          if (s < S) {
            av[j] = __ldg(a_row + (s >> 2));
            bv[j] = __ldg(b_row + (s >> 2));
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int s = base + j * 128 + lane * 4;
          if (s < S) {
            const float ab[4] = {av[j].x + bv[j].x, av[j].y + bv[j].y, av[j].z + bv[j].z, av[j].w + bv[j].w};
            // s is a multiple of 4: states s, s+2 are blanks, s+1, s+3 carry labels
            blank_acc += ex2_approx(ab[0] - lp_blank2 + loss2);
            if (s + 1 < S) {
              const int v = lab_s[s >> 1];
              const float x = ex2_approx(ab[1] - lp_row[v] + loss2);
              if (x != 0.f) atomicAdd(&my_bins[v], x);
            }
            if (s + 2 < S) blank_acc += ex2_approx(ab[2] - lp_blank2 + loss2);
            if (s + 3 < S) {
              const int v = lab_s[(s >> 1) + 1];
              const float x = ex2_approx(ab[3] - lp_row[v] + loss2);
              if (x != 0.f) atomicAdd(&my_bins[v], x);
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) blank_acc += __shfl_xor_sync(0xffffffffu, blank_acc, o);
      __syncwarp();
      float pv[2] = {0.f, 0.f}, dLdp[2] = {0.f, 0.f};
      float dot = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int v = lane + 32 * h;
        if (v < V) {
          const float occ = (v == blank) ? blank_acc : my_bins[v];
          pv[h] = pv_mine[h];
          const float g = expf(lp_mine[h]) - occ;
          dLdp[h] = g / (pv[h] + 1e-8f);
          dot += pv[h] * dLdp[h];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
      for (int h = 0; h < 2; ++h) dz[h] = pv[h] * (dLdp[h] - dot) * grad_scale;
      __syncwarp();
    }
    if (dz_f32 != nullptr) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (lane + 32 * h < V) dz_f32[ro * V + lane + 32 * h] = dz[h];
    }
    if (dz_packed != nullptr) {
      __nv_bfloat16* row = dz_packed + ro * row_elems;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (fp16) {  // (uniform)
          reinterpret_cast<uint16_t*>(row)[lane + 32 * h] = pack_16(dz[h], 1);
          continue;
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(dz[h]);
        row[lane + 32 * h] = hi;
        if (planes == 2) row[64 + lane + 32 * h] = __float2bfloat16_rn(dz[h] - __bfloat162float(hi));
      }
    }
  }
}

// Same gradient, without shared-memory atomics (the default).  The atomics of the kernel above
// (one per label state and frame, ~2 cycles per lane on the SM's single atomic unit) cost more than
// its HBM traffic.  Here the label positions of the utterance are counting-sorted by symbol once (by the
// spare warp of the alpha walker, see sort_labels_by_symbol): rank[i] = slot of label i in the
// symbol-sorted order, seg[v] = first slot of symbol v.  Per frame a warp then writes the occupancy
// term of every label state to its slot (plain conflict-light STS) and lane v adds up the
// contiguous segment of symbol v in label order — deterministic, no atomics.
struct GradCtx {
  const float* logp;
  const float* probs;
  __nv_bfloat16* dz_packed;
  float* dz_f32;
  float grad_scale;
  int T, V, blank, planes, fp16;
  const int* packed;  // smem [L_pad]: symbol | slot << 8 of every label position
  const int* seg;     // smem [VP + 1]: first slot of each symbol
};

// Normalisation of the occupancies  occ_t(v) = sum_{s: e[s]=v} alpha_t(s) beta_t(s) / (y_t(v) Z):
//   self_normalised = false  Z = exp(-loss[b]) from the finished alpha walk (two-launch path): offset2 = loss * log2 e;
//   self_normalised = true   (fused launch: the walks are still running) every frame is normalised by its own
//                            sum over all states — mathematically the same Z for every t — with the exponent
//                            offset taken from a max pass over the warp's first frame (the per-frame maxima of
//                            one utterance differ by at most log2 S).
struct GradNorm {
  bool self_normalised;
  bool have_offset;
  bool feasible;
  float offset2;  // added to every base-2 exponent
};

// The gradient row of ONE frame by ONE warp.  a_row / b_row: the frame's alpha / beta rows (base-2 logs), in
// global memory (L2-coherent loads: the fused launch reads rows written during the same launch) or already
// staged in shared memory (SMEM_ROWS).  lp_row [VP], xs [L_pad]: this warp's scratch.
// this lane's two symbols (lane, lane + 32) of a frame's log-probability and probability rows
struct FrameRows {
  float lp[2];
  float pv[2];
};
__device__ __forceinline__ FrameRows load_frame_rows(const GradCtx& c, int b, int t, int lane) {
  const size_t ro = static_cast<size_t>(b) * c.T + t;
  FrameRows r;
  r.lp[0] = c.logp[ro * VP + lane];
  r.lp[1] = c.logp[ro * VP + lane + 32];
  r.pv[0] = lane < c.V ? c.probs[ro * c.V + lane] : 0.f;
  r.pv[1] = lane + 32 < c.V ? c.probs[ro * c.V + lane + 32] : 0.f;
  return r;
}

template <int NP, bool SMEM_ROWS>  // NP: passes of 128 states whose loads are issued together
__device__ __forceinline__ void grad_frame(const GradCtx& c, int b, int t, int S, bool in_range, const FrameRows& fr,
                                           const float4* a_row, const float4* b_row, float* lp_row, float* xs,
                                           GradNorm& nrm, int lane) {
  const size_t ro = static_cast<size_t>(b) * c.T + t;
  const int V = c.V, blank = c.blank;
  const int* packed = c.packed;
  float dz[2] = {0.f, 0.f};  // symbols lane and lane + 32
  if (in_range && nrm.feasible) {
    const float lp_mine[2] = {fr.lp[0], fr.lp[1]};
    const float pv[2] = {fr.pv[0], fr.pv[1]};
    lp_row[lane] = lp_mine[0] * LOG2E;
    lp_row[lane + 32] = lp_mine[1] * LOG2E;
    __syncwarp();
    auto row4 = [](const float4* p) { return SMEM_ROWS ? *p : __ldcg(p); };
    if (!nrm.have_offset) {
      // max over the states of alpha + beta - lp: the exponent offset of this warp's frames
      float m = NEG;
      for (int s = lane * 4; s < S; s += 128) {
        const float4 av = row4(a_row + (s >> 2)), bv = row4(b_row + (s >> 2));
        const int2 pk = *reinterpret_cast<const int2*>(packed + (s >> 1));
        m = fmaxf(m, av.x + bv.x - lp_row[blank]);
        if (s + 1 < S) m = fmaxf(m, av.y + bv.y - lp_row[pk.x & 255]);
        if (s + 2 < S) m = fmaxf(m, av.z + bv.z - lp_row[blank]);
        if (s + 3 < S) m = fmaxf(m, av.w + bv.w - lp_row[pk.y & 255]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      nrm.offset2 = -m;
      nrm.have_offset = true;
      nrm.feasible = m > -1e29f;  // no alignment at all: zero gradient (the walk reports an infinite loss)
    }
    const float offset2 = nrm.offset2;
    const float off_blank = offset2 - lp_row[blank];
    float blank_acc = 0.f;
    for (int base = 0; base < S; base += 128 * NP) {
      // up to NP passes of 128 states issued together (2 NP x 16-byte loads in flight per lane)
      float4 av[NP], bv[NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const int s = base + j * 128 + lane * 4;
        if (s < S) {
          av[j] = row4(a_row + (s >> 2));
          bv[j] = row4(b_row + (s >> 2));
        }
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const int s = base + j * 128 + lane * 4;
        if (s < S) {
          // s is a multiple of 4: states s, s+2 are blanks, s+1, s+3 carry labels s/2, s/2+1
          // (one 8-byte read for both labels; past the last label it reads unused padding)
          const int2 pk = *reinterpret_cast<const int2*>(packed + (s >> 1));
          blank_acc += ex2_approx(av[j].x + bv[j].x + off_blank);
          if (s + 1 < S) xs[pk.x >> 8] = ex2_approx(av[j].y + bv[j].y - lp_row[pk.x & 255] + offset2);
          if (s + 2 < S) blank_acc += ex2_approx(av[j].z + bv[j].z + off_blank);
          if (s + 3 < S) xs[pk.y >> 8] = ex2_approx(av[j].w + bv[j].w - lp_row[pk.y & 255] + offset2);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) blank_acc += __shfl_xor_sync(0xffffffffu, blank_acc, o);
    __syncwarp();
    float occ[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int v = lane + 32 * h;
      if (v < V) {
        occ[h] = blank_acc;
        if (v != blank) {
          const int k0 = c.seg[v], k1 = c.seg[v + 1];
          float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
          int k = k0;
          for (; k + 4 <= k1; k += 4) {
            acc0 += xs[k];
            acc1 += xs[k + 1];
            acc2 += xs[k + 2];
            acc3 += xs[k + 3];
          }
          for (; k < k1; ++k) acc0 += xs[k];
          occ[h] = (acc0 + acc1) + (acc2 + acc3);
        }
      }
    }
    if (nrm.self_normalised) {
      float total = occ[0] + occ[1];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      const float inv = nrm.feasible && total > 0.f ? 1.0f / total : 0.f;
      occ[0] *= inv;
      occ[1] *= inv;
    }
    float dLdp[2] = {0.f, 0.f};
    float dot = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int v = lane + 32 * h;
      if (v < V) {
        const float g = expf(lp_mine[h]) - occ[h];
        dLdp[h] = g / (pv[h] + 1e-8f);
        dot += pv[h] * dLdp[h];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
    for (int h = 0; h < 2; ++h) dz[h] = nrm.feasible ? pv[h] * (dLdp[h] - dot) * c.grad_scale : 0.f;
    __syncwarp();
  }
  if (c.dz_f32 != nullptr) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (lane + 32 * h < V) c.dz_f32[ro * V + lane + 32 * h] = dz[h];
  }
  if (c.dz_packed != nullptr) {
    __nv_bfloat16* row = c.dz_packed + ro * (c.planes * 64);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (c.fp16) {  // (uniform)
        reinterpret_cast<uint16_t*>(row)[lane + 32 * h] = pack_16(dz[h], 1);
        continue;
      }
      const __nv_bfloat16 hi = __float2bfloat16_rn(dz[h]);
      row[lane + 32 * h] = hi;
      if (c.planes == 2) row[64 + lane + 32 * h] = __float2bfloat16_rn(dz[h] - __bfloat162float(hi));
    }
  }
}

// shared-memory carve-up of a gradient CTA: sorted labels, then per warp [lp_row VP | xs L_pad], then (staged
// rows only) per warp two [alpha row | beta row] buffers
struct GradSmem {
  int* packed;
  int* seg;
  float* per_warp;
  float* stage;
  int warp_stride;
};
__device__ __forceinline__ GradSmem grad_smem(uint8_t* smem_raw, int L_max, int nwarps) {
  const int L_pad = (L_max + 31) & ~31;
  GradSmem g;
  g.packed = reinterpret_cast<int*>(smem_raw);
  g.seg = g.packed + L_pad;
  g.per_warp = reinterpret_cast<float*>(g.seg + SORT_EXTRA);
  g.warp_stride = VP + L_pad;
  g.stage = g.per_warp + static_cast<size_t>(nwarps) * g.warp_stride;
  return g;
}
__device__ __forceinline__ void load_sorted_labels(const GradSmem& g, const int* __restrict__ sort_ws, int b, int L,
                                                   int L_max) {
  const int L_pad = (L_max + 31) & ~31;
  const int* src = sort_ws + static_cast<size_t>(b) * (L_pad + SORT_EXTRA);
  for (int i = threadIdx.x; i < L; i += blockDim.x) g.packed[i] = __ldcg(src + i);
  for (int i = threadIdx.x; i <= VP; i += blockDim.x) g.seg[i] = __ldcg(src + L_pad + i);
  __syncthreads();
}

// Two-launch path: one block per (utterance, chunk of frames), one warp per frame, after the lattice kernel.
// STAGED: the warp's next frame's alpha / beta rows travel to shared memory with cp.async (two row buffers per
// warp) while it works on the current frame — the kernel is bound by the latency of those rows, not by DRAM
// bandwidth (29 % of peak with direct loads), and registers cap it at three blocks per SM.
template <bool STAGED>
__global__ void ctc_grad_sorted_kernel(const float* __restrict__ logp, const float* __restrict__ probs,
                                       const int32_t* __restrict__ labels,
                                       const int32_t* __restrict__ input_len,
                                       const int32_t* __restrict__ label_len,
                                       const float* __restrict__ loss, const float* __restrict__ alpha,
                                       const float* __restrict__ beta, const int* __restrict__ sort_ws,
                                       __nv_bfloat16* __restrict__ dz_packed, float* __restrict__ dz_f32,
                                       float grad_scale, int T, int V, int L_max, int blank, int S_stride,
                                       int planes, int fp16, int frames_per_block) {
  extern __shared__ uint8_t smem_raw[];
  ptx::pdl_launch_dependents();
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int t_begin = blockIdx.x * frames_per_block;
  const int t_end = min(t_begin + frames_per_block, T);
  const int L = label_len[b];
  const int P = min(input_len[b], T);
  const int S = 2 * L + 1;
  ptx::pdl_wait();  // alpha / beta / loss / the sorted labels come from the lattice kernel right before
  const GradSmem g = grad_smem(smem_raw, L_max, nwarps);
  if (t_begin < P) load_sorted_labels(g, sort_ws, b, L, L_max);  // (block-uniform)
  const GradCtx c{logp, probs, dz_packed, dz_f32, grad_scale, T, V, blank, planes, fp16, g.packed, g.seg};
  const float loss_b = loss[b];
  GradNorm nrm{false, true, isfinite(loss_b), loss_b * LOG2E};
  float* lp_row = g.per_warp + warp * g.warp_stride;
  if constexpr (!STAGED) {
    for (int t = t_begin + warp; t < t_end; t += nwarps) {
      const size_t ro = static_cast<size_t>(b) * T + t;
      FrameRows fr = {};
      if (t < P) fr = load_frame_rows(c, b, t, lane);
      grad_frame<4, false>(c, b, t, S, t < P, fr, reinterpret_cast<const float4*>(alpha + ro * S_stride),
                           reinterpret_cast<const float4*>(beta + ro * S_stride), lp_row, lp_row + VP, nrm, lane);
    }
  } else {
    float* stage = g.stage + static_cast<size_t>(warp) * 4 * S_stride;
    auto prefetch_rows = [&](int t, int buf) {
      const float* a = alpha + (static_cast<size_t>(b) * T + t) * S_stride;
      const float* bb = beta + (static_cast<size_t>(b) * T + t) * S_stride;
      float* dst = stage + buf * 2 * S_stride;
      for (int i = lane * 4; i < S; i += 128) {
        cp_async16_cg(dst + i, a + i);
        cp_async16_cg(dst + S_stride + i, bb + i);
      }
      cp_async_commit();
    };
    const int t_last = min(t_end, P);  // frames beyond P only get zero rows
    int buf = 0;
    FrameRows fr = {}, fr_next = {};
    if (t_begin + warp < t_last && nrm.feasible) {
      prefetch_rows(t_begin + warp, 0);
      fr = load_frame_rows(c, b, t_begin + warp, lane);
    }
    for (int t = t_begin + warp; t < t_end; t += nwarps) {
      const bool live = t < t_last && nrm.feasible;
      if (live) {
        if (t + nwarps < t_last) {
          prefetch_rows(t + nwarps, buf ^ 1);
          fr_next = load_frame_rows(c, b, t + nwarps, lane);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
      }
      grad_frame<1, true>(c, b, t, S, t < P, fr, reinterpret_cast<const float4*>(stage + buf * 2 * S_stride),
                          reinterpret_cast<const float4*>(stage + buf * 2 * S_stride + S_stride), lp_row, lp_row + VP, nrm,
                          lane);
      if (live) {
        __syncwarp();  // every lane is done with this buffer before the prefetch two frames on refills it
        fr = fr_next;
        buf ^= 1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused launch: lattices AND gradient in one grid.  CTAs [0, n_walk) are the lattice walkers above; the
// remaining CTAs are persistent GRADIENT CTAs.  The gradient of frame t needs alpha_t (ready once the
// alpha walk has passed t) and beta_t (ready once the beta walk, coming from the other end, has passed
// t): the frames in the middle of an utterance become available when both walks are half way, the
// outermost ones when they finish.  So the bandwidth-bound gradient phase runs underneath the second
// half of the latency-bound walks — on the SMs' idle issue slots and against lattice rows that are still
// in L2 — instead of as a second launch after them.
//
// Gradient CTA g serves utterance g % B together with the other CTAs of that residue: their warps deal
// the utterance's frames round-robin IN ORDER OF AVAILABILITY (the middle frame first, then alternately
// one further down / up).  Every warp is its own pipeline — no block barrier after the label load: lane 0
// polls the two progress flags of the utterance (ld.acquire.gpu), the frame's rows are copied to shared
// memory with cp.async (L2 path), and while they are in flight the warp works on the frame it fetched
// before (two row buffers per warp).  Walkers never wait for gradient CTAs, and every CTA of the grid is
// co-resident (the launcher sizes the grid with the occupancy API), so the waits cannot deadlock; they
// are bounded anyway.
template <bool STAGED>
__device__ __noinline__ void grad_stream(uint8_t* smem_raw, int g_index, int G, const float* __restrict__ logp,
                                         const float* __restrict__ probs, const int32_t* __restrict__ input_len,
                                         const int32_t* __restrict__ label_len, const float* __restrict__ alpha,
                                         const float* __restrict__ beta, const int* __restrict__ sort_ws,
                                         const int* __restrict__ progress, __nv_bfloat16* __restrict__ dz_packed,
                                         float* __restrict__ dz_f32, float grad_scale, int B, int T, int V,
                                         int L_max, int blank, int S_stride, int planes, int fp16, int dbg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int b = g_index % B;
  if (b >= B || g_index >= G) return;
  const int q = g_index / B;                 // this CTA's index among the CTAs serving utterance b
  const int Q = (G - 1 - b) / B + 1;         // ... and their number
  const int L = label_len[b];
  const int P = min(input_len[b], T);
  const int S = 2 * L + 1;
  const int* flag_a = progress + 2 * b * FLAG_STRIDE;
  const int* flag_b = flag_a + FLAG_STRIDE;
  // lane 0 only.  The flags only grow: the values seen last are kept (per warp, and per CTA in shared memory:
  // a warp first looks at what its neighbours have already seen), and memory is asked again only when they do
  // not cover the frame yet — a ld.acquire.gpu is an L2 round trip of ~1 us, and thousands of warps hammering
  // the same flag line delay the walkers' updates of it.  `lag`: steps the slower walk is still short of.
  const GradSmem gs = grad_smem(smem_raw, L_max, nwarps);
  volatile int* seen_cta = reinterpret_cast<volatile int*>(gs.seg + VP + 2);  // [2] (inside SORT_EXTRA's padding)
  int seen_a = 0, seen_b = 0, lag = 0;
  auto available = [&](int t) {
    const int need_a = t + 1, need_b = P - t;
    if (seen_a < need_a) seen_a = max(seen_a, seen_cta[0]);
    if (seen_b < need_b) seen_b = max(seen_b, seen_cta[1]);
    if (seen_a < need_a) {
      seen_a = ld_acquire_gpu(flag_a);
      seen_cta[0] = seen_a;  // (racy maximum: a stale smaller value only costs a neighbour one more poll)
    }
    if (seen_a >= need_a && seen_b < need_b) {
      seen_b = ld_acquire_gpu(flag_b);
      seen_cta[1] = seen_b;
    }
    lag = max(need_a - seen_a, need_b - seen_b);
    return lag <= 0;
  };
  auto wait_available = [&](int t) {  // whole warp; returns once frame t's rows are visible to every lane
    if (dbg == 3) return;  // (measurement aid: no waiting — results are garbage)
    if (lane == 0) {
      unsigned spins = 0;
      while (!available(t)) {
        // a lattice step takes ~0.1 us: sleep about half of what is still missing, at most 4 us
        __nanosleep(min(4000, max(100, lag * 50)));
        if (++spins > (1u << 22)) {
          printf("speechless_b200: CTC gradient warp timed out waiting for the lattice walk (utterance %d)\n", b);
          __trap();
        }
      }
    }
    __syncwarp();
  };
  // the sorted labels are published together with the first alpha progress
  if (threadIdx.x == 0) {
    seen_cta[0] = 0;
    seen_cta[1] = 0;
  }
  if (P > 0) {
    if (threadIdx.x == 0 && dbg != 3) {
      unsigned spins = 0;
      while (ld_acquire_gpu(flag_a) < 1) {
        __nanosleep(500);
        if (++spins > (1u << 22)) __trap();
      }
    }
    __syncthreads();
  }
  if (P > 0) load_sorted_labels(gs, sort_ws, b, L, L_max);
  __syncthreads();
  if (dbg == 2) return;
  const GradCtx c{logp, probs, dz_packed, dz_f32, grad_scale, T, V, blank, planes, fp16, gs.packed, gs.seg};
  GradNorm nrm{true, false, true, 0.f};
  float* lp_row = gs.per_warp + warp * gs.warp_stride;
  float* xs = lp_row + VP;
  float* stage = gs.stage + static_cast<size_t>(warp) * 4 * S_stride;

  // i-th frame of the utterance in order of availability: the middle one, then alternately one further
  // down / up, then what is left of the longer (upper) side; frames >= P (zero rows) keep their index
  const int mid = P / 2;
  auto frame_of = [&](int i) {
    if (i >= P) return i;
    if (i < 2 * mid) return (i & 1) ? mid - ((i + 1) >> 1) : mid + (i >> 1);
    return i;  // i == 2 * mid == P - 1 (odd P): the last frame
  };
  auto prefetch_rows = [&](int t, int buf) {
    const float* a = alpha + (static_cast<size_t>(b) * T + t) * S_stride;
    const float* bb = beta + (static_cast<size_t>(b) * T + t) * S_stride;
    float* dst = stage + buf * 2 * S_stride;
    for (int i = lane * 4; i < S; i += 128) {
      cp_async16_cg(dst + i, a + i);
      cp_async16_cg(dst + S_stride + i, bb + i);
    }
    cp_async_commit();
  };
  const int stride = Q * nwarps;
  int buf = 0;
  int i = q * nwarps + warp;
  FrameRows fr = {}, fr_next = {};  // the frame's log-prob / prob rows travel one frame ahead in registers
  if (i < P) {
    wait_available(frame_of(i));
    if (STAGED) prefetch_rows(frame_of(i), 0);
    fr = load_frame_rows(c, b, frame_of(i), lane);
  }
  for (; i < T; i += stride) {
    const int t = frame_of(i);
    const int i_next = i + stride;
    bool next_issued = false;
    if (i < P) {
      if (i_next < P) {
        // fetch the next frame's rows now if the walks are already past it
        const int t_next = frame_of(i_next);
        int ok = 0;
        if (lane == 0) ok = (dbg == 3 || available(t_next)) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          if (STAGED) prefetch_rows(t_next, buf ^ 1);
          fr_next = load_frame_rows(c, b, t_next, lane);
          next_issued = true;
        }
      }
      if (STAGED) {
        if (next_issued)
          cp_async_wait<1>();
        else
          cp_async_wait<0>();
        __syncwarp();
      }
    }
    const size_t ro = static_cast<size_t>(b) * T + t;
    const float* a_row = STAGED ? stage + buf * 2 * S_stride : alpha + ro * S_stride;
    const float* b_row = STAGED ? stage + buf * 2 * S_stride + S_stride : beta + ro * S_stride;
    grad_frame<1, STAGED>(c, b, t, S, i < P, fr, reinterpret_cast<const float4*>(a_row),
                          reinterpret_cast<const float4*>(b_row), lp_row, xs, nrm, lane);
    if (i_next < P) {
      __syncwarp();  // every lane is done with the buffers
      if (!next_issued) {
        wait_available(frame_of(i_next));
        if (STAGED) prefetch_rows(frame_of(i_next), buf ^ 1);
        fr_next = load_frame_rows(c, b, frame_of(i_next), lane);
      }
      fr = fr_next;
      buf ^= 1;
    }
  }
}

template <int SPT, int K, bool CL>
__global__ void __launch_bounds__(1024)
    ctc_fused_kernel(const float* __restrict__ logp, const float* __restrict__ probs,
                     const int32_t* __restrict__ labels, const int32_t* __restrict__ input_len,
                     const int32_t* __restrict__ label_len, float* __restrict__ loss,
                     float* __restrict__ beta_loss, float* __restrict__ alpha, float* __restrict__ beta,
                     int* __restrict__ sort_ws, int* __restrict__ progress, __nv_bfloat16* __restrict__ dz_packed,
                     float* __restrict__ dz_f32, float grad_scale, int B, int T, int V, int L_max, int blank,
                     int S_stride, int col_stride, int planes, int fp16, int n_walk, int walkers_per_sm,
                     int staged, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  ptx::pdl_launch_dependents();
  // Roles.  The walks are latency bound: two walkers on one SM share its four schedulers and both slow down
  // (measured: the same walkers inside a 592-CTA grid took 82 us instead of 65 when roles went by block index,
  // because the block scheduler does not spread consecutive CTAs one per SM once several fit).  So a CTA
  // becomes a walker if it is among the first `walkers_per_sm` CTAs to arrive on ITS SM (%smid) and walk
  // units are left; everything else computes gradients.  Every unit is claimed whatever the placement: the last
  // n_walk CTAs to arrive also try to claim.  The counters sit behind the progress flags and are zeroed with
  // them.  Clusters keep roles by block index.
  int walk_unit = -1, g = 0;
  if constexpr (CL) {
    if (static_cast<int>(blockIdx.x) < n_walk)
      walk_unit = static_cast<int>(blockIdx.x) / static_cast<int>(ptx::cluster_nctarank());
    else
      g = static_cast<int>(blockIdx.x) - n_walk;
  } else {
    __shared__ int role_s[2];
    if (threadIdx.x == 0) {
      // counters behind the flags, one 128-byte line each (atomics on one line serialise in its L2 slice):
      // [0] walk units claimed, [32] gradient CTAs, [64] CTAs arrived, [96 + smid] CTAs arrived on that SM
      // (plain fetch-and-adds: a compare-and-swap loop of 600 CTAs on one word took longer than the walk)
      int* counters = progress + 2 * B * FLAG_STRIDE;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const int slot = atomicAdd(counters + 96 + static_cast<int>(smid & 255u), 1);
      const int ticket = atomicAdd(counters + 64, 1);
      // claim a walk unit if this CTA is among the first on its SM — or among the last n_walk CTAs to ARRIVE,
      // which mop up whatever an unexpected placement has left unclaimed (each of them tries exactly once, so
      // every unit finds a walker whatever the hardware does; going by block index instead of arrival made
      // the high-index CTAs, which start at the same time as everybody else, double up on SMs)
      int unit = -1;
      if (slot < walkers_per_sm || ticket >= static_cast<int>(gridDim.x) - n_walk) {
        const int u = atomicAdd(counters, 1);
        if (u < n_walk) unit = u;
      }
      role_s[0] = unit;
      role_s[1] = unit < 0 ? atomicAdd(counters + 32, 1) : 0;
    }
    __syncthreads();
    walk_unit = role_s[0];
    g = role_s[1];
  }
  if (walk_unit >= 0) {
    lattice_cta_body<SPT, K, CL>(smem_raw, walk_unit, logp, labels, input_len, label_len, loss, beta_loss, alpha, beta,
                                 sort_ws, T, L_max, blank, S_stride, col_stride, dbg == 4 ? nullptr : progress, dbg);
    return;
  }
  if (dbg == 1 || dbg >= 4) return;  // measurement aids (SL_CTC_DBG): walkers (+ publication) only
  ptx::pdl_wait();  // logp / probs come from the output_conv kernel right before
  const int G = static_cast<int>(gridDim.x) - n_walk;
  if (staged)
    grad_stream<true>(smem_raw, g, G, logp, probs, input_len, label_len, alpha, beta, sort_ws, progress, dz_packed,
                      dz_f32, grad_scale, B, T, V, L_max, blank, S_stride, planes, fp16, dbg);
  else
    grad_stream<false>(smem_raw, g, G, logp, probs, input_len, label_len, alpha, beta, sort_ws, progress, dz_packed,
                       dz_f32, grad_scale, B, T, V, L_max, blank, S_stride, planes, fp16, dbg);
}

// Greedy decode: one warp per utterance, 32 frames per iteration.  argmax (lowest index
// wins ties), emit iff not blank and (merge_repeated ? differs from previous frame's
// argmax : true); positions by ballot/popc prefix.
__global__ void ctc_greedy_kernel(const float* __restrict__ probs,
                                  const int32_t* __restrict__ input_len, int32_t* __restrict__ out,
                                  int32_t* __restrict__ out_len, int B, int T, int V, int blank,
                                  int merge_repeated) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int P = min(input_len[b], T);
  const float* pb = probs + static_cast<size_t>(b) * T * V;
  int32_t* ob = out + static_cast<size_t>(b) * T;
  int count = 0;
  int carry = -1;  // argmax of the frame before this 32-frame group
  for (int t0 = 0; t0 < P; t0 += 32) {
    const int t = t0 + lane;
    int c = -1;
    if (t < P) {
      const float* row = pb + static_cast<size_t>(t) * V;
      float best = row[0];
      c = 0;
      for (int v = 1; v < V; ++v) {
        const float x = row[v];
        if (x > best) {
          best = x;
          c = v;
        }
      }
    }
    int prev = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) prev = carry;
    const bool emit = (t < P) && (c != blank) && !(merge_repeated && c == prev);
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    if (emit) ob[count + __popc(m & ((1u << lane) - 1))] = c;
    count += __popc(m);
    carry = __shfl_sync(0xffffffffu, c, 31);
  }
  for (int i = count + lane; i < T; i += 32) ob[i] = -1;
  if (lane == 0) out_len[b] = count;
}

}  // namespace

int ctc_s_stride(int L_max) { return ((2 * L_max + 1) + 31) & ~31; }

size_t ctc_workspace_bytes(int B, int T, int L_max) {
  const size_t lat = static_cast<size_t>(B) * T * ctc_s_stride(L_max) * sizeof(float);
  // alpha | beta | beta_loss[B] (256-byte slot granularity) | per utterance: symbol-sorted label slots + segments
  const size_t bl = (static_cast<size_t>(B) * sizeof(float) + 255) & ~static_cast<size_t>(255);
  const size_t sort = static_cast<size_t>(B) * (((L_max + 31) & ~31) + SORT_EXTRA) * sizeof(int);
  // ... | progress flags of the fused launch: steps completed per (utterance, direction)
  const size_t flags = (static_cast<size_t>(2 * B) * FLAG_STRIDE + ROLE_COUNTERS) * sizeof(int);
  return 2 * lat + bl + sort + flags + 512;
}

// co-resident CTAs of a fused-launch configuration on the current device, cached (benign race: idempotent)
static cudaError_t fused_capacity(const void* kern, int threads, size_t smem, int cluster, int sms, int* capacity) {
  struct Entry {
    const void* kern;
    int threads, cluster, dev, capacity;
    size_t smem;
  };
  static Entry cache[64];
  static int used = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  for (int i = 0; i < used; ++i)
    if (cache[i].kern == kern && cache[i].threads == threads && cache[i].smem == smem && cache[i].cluster == cluster &&
        cache[i].dev == dev) {
      *capacity = cache[i].capacity;
      return cudaSuccess;
    }
  if (smem > 40 * 1024) {  // (the kernel also has a few bytes of static shared memory: opt in below the 48 KB line)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  int cap = 0;
  if (cluster > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / cluster * cluster);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      clusters = 0;
    }
    cap = clusters * cluster;
  } else {
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    cap = per_sm * sms;
  }
  if (used < 64) cache[used++] = Entry{kern, threads, cluster, dev, cap, smem};
  *capacity = cap;
  return cudaSuccess;
}

int ctc_loss_launch(const float* logp, const float* probs, const int32_t* labels,
                    const int32_t* input_len, const int32_t* label_len, float* loss,
                    void* dlogits_packed, float* dlogits_f32, float grad_scale, int B, int T, int V,
                    int L_max, int blank, int prec, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream) {
  const int planes = prec_planes(prec), fp16 = prec_fp16(prec);
  SL_REQUIRE(V <= VP && V >= 2, "CTC kernels support 2..64 symbols (incl. blank)");
  SL_REQUIRE(blank >= 0 && blank < V, "blank out of range");
  SL_REQUIRE(L_max >= 1, "L_max must be >= 1 (pad empty label batches to width 1)");
  SL_REQUIRE(workspace_bytes >= ctc_workspace_bytes(B, T, L_max), "CTC workspace too small");
  const int S_stride = ctc_s_stride(L_max);
  const size_t lat = static_cast<size_t>(B) * T * S_stride;
  float* alpha = reinterpret_cast<float*>(workspace);
  float* beta = alpha + lat;
  float* beta_loss = beta + lat;
  int* sort_ws = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(beta_loss) +
                                        ((static_cast<size_t>(B) * sizeof(float) + 255) & ~static_cast<size_t>(255)));
  const size_t sort_bytes = static_cast<size_t>(B) * (((L_max + 31) & ~31) + SORT_EXTRA) * sizeof(int);
  int* progress = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(sort_ws) + ((sort_bytes + 255) & ~static_cast<size_t>(255)));
  const bool want_grad = dlogits_packed != nullptr || dlogits_f32 != nullptr;
  if (want_grad) SL_REQUIRE(probs != nullptr, "gradient needs the softmax probabilities");
  bool grad_done = false;

  const int S_max = 2 * L_max + 1;
  SL_REQUIRE(S_max <= 4096, "label too long (max 2047 characters)");
  const char* legacy_env = std::getenv("SL_CTC_LEGACY");  // 1: one barrier per step, 2: wavefront (experiments)
  const int legacy = legacy_env ? std::atoi(legacy_env) : 0;
  if (legacy == 0) {
    // halo-blocked kernel: (states per lane, steps per barrier); owned states per warp = 32*SPT - 2K
    int spt = S_max <= 8 * 48 ? 2 : 4;
    int kk = 8;
    const char* spt_env = std::getenv("SL_CTC_SPT");  // tuning aids
    if (spt_env) spt = std::atoi(spt_env);
    if (const char* e = std::getenv("SL_CTC_K")) kk = std::atoi(e);
    int own = 0, nw = 0, cluster = 1;
    for (;;) {
      own = 32 * spt - 2 * kk;
      SL_REQUIRE(own > 0, "SL_CTC_SPT / SL_CTC_K combination not built");
      const int nw_all = (S_max + own - 1) / own;
      // few long utterances: split the warps of one lattice over a cluster of CTAs (halo exchange through
      // DSMEM, cluster barrier per K-block) so that the whole GPU works on the recurrence
      // (a cluster barrier costs ~1500 cycles against ~50 for a CTA barrier — measured — so clusters pay only
      // for lattices of more than a few warps, and run with K = 32: 16 x 60 s, S = 1801 on one B200 took
      // 1.32 ms without clusters, 0.77 / 0.58 / 0.48 ms with clusters of 4 and K = 8 / 16 / 32)
      cluster = 1;
      if (spt >= 4)
        while (cluster < 8 && 2 * B * cluster * 2 <= 148 && nw_all >= 4 * cluster) cluster *= 2;
      if (const char* e = std::getenv("SL_CTC_CLUSTER")) cluster = std::max(1, std::atoi(e));
      if (cluster > 1 && kk == 8 && !std::getenv("SL_CTC_K")) {
        kk = 32;
        continue;
      }
      SL_REQUIRE(cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8, "SL_CTC_CLUSTER must be 1, 2, 4 or 8");
      nw = (nw_all + cluster - 1) / cluster;  // compute warps per CTA (+ the label-sorter warp <= 32)
      if (nw <= 31 || spt_env || spt >= 8) break;
      spt *= 2;
    }
    SL_REQUIRE(nw <= 31, "SL_CTC_SPT too small for this label length");
    const int col_stride = ((nw * own + 2 * kk) + 3) & ~3;
    size_t smem = (2 * CHUNK * VP + 2 * col_stride + VP + 8) * sizeof(float);
    if (const char* e = std::getenv("SL_CTC_EXTRA_SMEM")) smem += static_cast<size_t>(std::atoi(e));  // measurement aid
    bool launched = false;
    // Fused launch (opt-in, SL_CTC_FUSED=1): gradient CTAs ride along with the walkers (ctc_fused_kernel);
    // SL_CTC_GRAD_CTAS bounds their number.  Measured on B200 at the bench shape (B = 64, 626 frames, S = 301):
    // the gradient phase does hide under the second half of the walks (8 us left over instead of 42), but the
    // same walkers run 15-20 % slower inside the big co-resident grid (65 -> 79 us with every gradient CTA
    // exiting at once), so the launch pair and the fused launch both come to 0.107 ms; with clusters (long-form
    // shapes) the fused launch is slower (0.93 vs 0.65 ms).  Hence two launches by default.
    const char* fused_env = std::getenv("SL_CTC_FUSED");
    if (want_grad && fused_env && std::atoi(fused_env) != 0) {
      const int threads = (nw + 1) * 32;
      const int nwarps = nw + 1;
      const int L_pad = (L_max + 31) & ~31;
      size_t gsmem = (L_pad + SORT_EXTRA) * sizeof(int) + static_cast<size_t>(nwarps) * (VP + L_pad) * sizeof(float);
      // row staging (two alpha + beta row pairs per warp) while a gradient CTA stays within a quarter of the
      // SM's shared memory, i.e. while it does not lower the number of co-resident CTAs
      const size_t stage_bytes = static_cast<size_t>(nwarps) * 4 * S_stride * sizeof(float);
      int staged = gsmem + stage_bytes <= 56 * 1024 ? 1 : 0;
      if (const char* e = std::getenv("SL_CTC_GRAD_STAGED")) staged = std::atoi(e) != 0 && gsmem + stage_bytes <= 200 * 1024;
      if (staged) gsmem += stage_bytes;
      const size_t fsmem = smem > gsmem ? smem : gsmem;
      const int n_walk = 2 * B * cluster;
      // no more gradient warps than frames; at least one gradient CTA per utterance (else: two launches)
      const int max_useful = B * ((T + nwarps - 1) / nwarps);
      int dev = 0, sms = 148;
      SL_CUDA(cudaGetDevice(&dev));
      SL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      const char* dbg_env = std::getenv("SL_CTC_DBG");  // measurement aids, see ctc_fused_kernel
      const int dbg = dbg_env ? std::atoi(dbg_env) : 0;
      int want_ctas = 1 << 20;  // as many as can be co-resident: the gradient phase is latency bound
      if (const char* e = std::getenv("SL_CTC_GRAD_CTAS")) want_ctas = std::max(1, std::atoi(e));
#define SL_LAUNCH_FUSED(SPT, KK)                                                                    \
  if (!launched && spt == SPT && kk == KK) {                                                        \
    auto kern = cluster > 1 ? ctc_fused_kernel<SPT, KK, true> : ctc_fused_kernel<SPT, KK, false>;   \
    /* every CTA of the grid must be co-resident: the gradient CTAs wait for the walkers.  The occupancy  \
       query (and the opt-in shared-memory size) are cached per (kernel, block, smem): they cost more host \
       time than the kernel runs */                                                                  \
    int capacity = 0;                                                                               \
    SL_CUDA(fused_capacity(reinterpret_cast<const void*>(kern), threads, fsmem, cluster, sms, &capacity)); \
    int n_grad = std::min(std::min(want_ctas, max_useful), capacity - n_walk);                     \
    n_grad = n_grad / cluster * cluster;                                                            \
    if (std::getenv("SL_CTC_VERBOSE"))                                                              \
      fprintf(stderr, "ctc fused: threads %d smem %zu capacity %d n_walk %d n_grad %d staged %d\n", threads, fsmem, \
              capacity, n_walk, n_grad, staged);                                                    \
    if (n_grad >= B) {                                                                              \
      SL_CUDA(cudaMemsetAsync(progress, 0, (static_cast<size_t>(2 * B) * FLAG_STRIDE + ROLE_COUNTERS) * sizeof(int), \
                              stream));                                                             \
      SL_CUDA(launch_pdl_cluster(PDL_CTC, ClusterX{cluster}, kern, dim3(n_walk + n_grad), dim3(threads), fsmem, stream, \
                                 logp, probs, labels, input_len, label_len, loss, beta_loss, alpha, beta, sort_ws,  \
                                 progress, reinterpret_cast<__nv_bfloat16*>(dlogits_packed), dlogits_f32, grad_scale, \
                                 B, T, V, L_max, blank, S_stride, col_stride, planes, fp16, n_walk,  \
                                 (n_walk + sms - 1) / sms, staged, dbg));                           \
      launched = true;                                                                              \
      grad_done = true;                                                                             \
    }                                                                                               \
  }
      SL_LAUNCH_FUSED(2, 4)
      SL_LAUNCH_FUSED(2, 8)
      SL_LAUNCH_FUSED(2, 16)
      SL_LAUNCH_FUSED(4, 8)
      SL_LAUNCH_FUSED(4, 16)
      SL_LAUNCH_FUSED(4, 32)
      SL_LAUNCH_FUSED(8, 32)
      SL_LAUNCH_FUSED(8, 8)
      SL_LAUNCH_FUSED(8, 16)
#undef SL_LAUNCH_FUSED
    }
#define SL_LAUNCH_HALO(SPT, KK)                                                                     \
  if (!launched && spt == SPT && kk == KK) {                                                        \
    auto kern = cluster > 1 ? ctc_lattice_halo_kernel<SPT, KK, true> : ctc_lattice_halo_kernel<SPT, KK, false>; \
    if (smem > 48 * 1024)                                                                           \
      SL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    SL_CUDA(launch_pdl_cluster(PDL_CTC, ClusterX{cluster}, kern, dim3(2 * B * cluster), dim3((nw + 1) * 32), smem, \
                               stream, logp, labels, input_len, label_len, loss, beta_loss, alpha, beta, sort_ws, T, \
                               L_max, blank, S_stride, col_stride));                                \
    launched = true;                                                                                \
  }
    SL_LAUNCH_HALO(2, 4)
    SL_LAUNCH_HALO(2, 8)
    SL_LAUNCH_HALO(2, 16)
    SL_LAUNCH_HALO(4, 8)
    SL_LAUNCH_HALO(4, 16)
    SL_LAUNCH_HALO(4, 32)
    SL_LAUNCH_HALO(8, 32)
    SL_LAUNCH_HALO(8, 8)
    SL_LAUNCH_HALO(8, 16)
#undef SL_LAUNCH_HALO
    SL_REQUIRE(launched, "SL_CTC_SPT / SL_CTC_K combination not built");
    SL_CUDA(cudaGetLastError());
  } else {
  // one state per thread while the lattice fits a CTA (measured: 0.149 / 0.166 / 0.176 ms for 1 / 2 / 4
  // states per thread at S = 301: more warps hide the dependent-instruction latency better)
  int spt = S_max <= 1024 ? 1 : 2;
  if (S_max > 2048) spt = 4;
  int threads = ((S_max + spt - 1) / spt + 31) & ~31;
  if (threads < 64) threads = 64;
  const int wspt = S_max <= 31 * 64 ? 2 : 4;  // wavefront kernel: 2 states per lane (4 beyond 1984 states)
  const int n_compute = (S_max + 32 * wspt - 1) / (32 * wspt);
  // Measured on B200 (B=64, P=625, S=301): the wavefront kernel needs 0.21 ms, the block-barrier
  // kernel 0.11 ms — a lone warp per scheduler executes its ~160 dependent instructions per
  // step at ~4 cycles each, while the barrier version interleaves warps.
  if (legacy == 2 && n_compute <= 31) {  // + one loader warp <= 1024 threads
    const int wthreads = (n_compute + 1) * 32;
    const size_t wsmem = LP_RING * VP * sizeof(float) + static_cast<size_t>(n_compute) * EDGE_RING * sizeof(uint4) +
                         static_cast<size_t>(n_compute) * 32 * wspt * sizeof(float) + (n_compute + 1) * sizeof(int);
    if (wspt == 2)
      ctc_alpha_beta_wave_kernel<2><<<2 * B, wthreads, wsmem, stream>>>(logp, labels, input_len, label_len, loss,
                                                                        beta_loss, alpha, beta, T, L_max, blank,
                                                                        S_stride);
    else
      ctc_alpha_beta_wave_kernel<4><<<2 * B, wthreads, wsmem, stream>>>(logp, labels, input_len, label_len, loss,
                                                                        beta_loss, alpha, beta, T, L_max, blank,
                                                                        S_stride);
    SL_CUDA(cudaGetLastError());
  } else {
  const int col_stride = (threads * spt > S_stride ? threads * spt : S_stride) + 2;
  const size_t smem = (2 * CHUNK * VP + 2 * col_stride) * sizeof(float) + S_stride * sizeof(int);
#define SL_LAUNCH_AB(SPT)                                                                          \
  do {                                                                                             \
    if (smem > 48 * 1024)                                                                          \
      SL_CUDA(cudaFuncSetAttribute(ctc_alpha_beta_kernel<SPT>,                                     \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    SL_CUDA(launch_pdl(PDL_CTC, ctc_alpha_beta_kernel<SPT>, dim3(2 * B), dim3(threads), smem, stream, logp, labels, \
                       input_len, label_len, loss, beta_loss, alpha, beta, T, L_max, blank, S_stride,  \
                       col_stride));                                                               \
  } while (0)
  if (spt == 1)
    SL_LAUNCH_AB(1);
  else if (spt == 2)
    SL_LAUNCH_AB(2);
  else
    SL_LAUNCH_AB(4);
#undef SL_LAUNCH_AB
  SL_CUDA(cudaGetLastError());
  }
  }

  if (want_grad && !grad_done) {
    // frames per block (8 warps): measured at the bench shape, loss + gradient: direct loads 16 / 32 / 64 frames
    // -> 0.1068 / 0.1079 / 0.1154 ms; rows staged with cp.async one frame ahead -> 0.1015 / 0.1006 / 0.0985 ms
    int frames_per_block = 64;
    bool fpb_given = false;
    if (const char* e = std::getenv("SL_CTC_GRAD_FPB")) {  // tuning aid
      frames_per_block = std::max(1, std::atoi(e));
      fpb_given = true;
    }
    const int warps = 8;
    dim3 grid((T + frames_per_block - 1) / frames_per_block, B);
    const int L_pad = (L_max + 31) & ~31;
    if (legacy == 0) {
      size_t gsmem = (L_pad + SORT_EXTRA) * sizeof(int) + static_cast<size_t>(warps) * (VP + L_pad) * sizeof(float);
      const size_t stage_bytes = static_cast<size_t>(warps) * 4 * S_stride * sizeof(float);
      bool staged = gsmem + stage_bytes <= 64 * 1024 && frames_per_block > warps;  // (>= 2 frames per warp to overlap)
      if (const char* e = std::getenv("SL_CTC_GRAD_STAGED")) staged = std::atoi(e) != 0 && gsmem + stage_bytes <= 200 * 1024;
      if (staged) gsmem += stage_bytes;
      if (!staged && !fpb_given) {  // direct loads like fewer frames per block
        frames_per_block = 32;
        grid = dim3((T + frames_per_block - 1) / frames_per_block, B);
      }
      auto grad_kernel = staged ? ctc_grad_sorted_kernel<true> : ctc_grad_sorted_kernel<false>;
      static size_t opted_in[2] = {0, 0};  // largest dynamic shared memory size set so far, per variant
      if (gsmem > 48 * 1024 && gsmem > opted_in[staged ? 1 : 0]) {
        SL_CUDA(cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(gsmem)));
        opted_in[staged ? 1 : 0] = gsmem;
      }
      SL_CUDA(launch_pdl(PDL_CTC, grad_kernel, grid, dim3(warps * 32), gsmem, stream, logp, probs, labels,
                         input_len, label_len, static_cast<const float*>(loss), static_cast<const float*>(alpha),
                         static_cast<const float*>(beta), static_cast<const int*>(sort_ws),
                         reinterpret_cast<__nv_bfloat16*>(dlogits_packed), dlogits_f32, grad_scale, T, V, L_max, blank,
                         S_stride, planes, fp16, frames_per_block));
    } else {
      const size_t gsmem = L_pad * sizeof(int) + warps * 2 * VP * sizeof(float);
      SL_CUDA(launch_pdl(PDL_CTC, ctc_grad_kernel, grid, dim3(warps * 32), gsmem, stream, logp, probs, labels,
                         input_len, label_len, static_cast<const float*>(loss), static_cast<const float*>(alpha),
                         static_cast<const float*>(beta), reinterpret_cast<__nv_bfloat16*>(dlogits_packed),
                         dlogits_f32, grad_scale, T, V, L_max, blank, S_stride, planes, fp16, frames_per_block));
    }
  }
  return 0;
}

int ctc_greedy_launch(const float* probs, const int32_t* input_len, int32_t* out, int32_t* out_len,
                      int B, int T, int V, int blank, int merge_repeated, cudaStream_t stream) {
  const int warps = 4;
  ctc_greedy_kernel<<<(B + warps - 1) / warps, warps * 32, 0, stream>>>(
      probs, input_len, out, out_len, B, T, V, blank, merge_repeated);
  SL_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sl
