// conv_umma.cu — Conv1D forward and input-gradient as an sm_100a implicit GEMM.
//
// Replaces keras.layers.Conv1D(padding="same") forward (reference net.py:304-305) and
// the TF autodiff input gradient of the same layer.  Nothing is materialised as im2col:
// with channels-last activations (B, T, C) the A operand of tap j is the *same* tile of
// 128 time frames shifted by j rows, so one TMA box load per (tap, 64-channel chunk)
// with signed coordinates fetches it, and TMA's out-of-bounds zero fill implements the
// TF "SAME" padding (asymmetric pads included) for free.
//
//   D[128 frames x BN filters] (TMEM, fp32) += A[128 x 64] (smem, K-major, SW128)
//                                              * B[BN x 64]^T (smem, K-major, SW128)
//
// Warp roles (256 threads, 1 CTA / SM, persistent over output tiles):
//   warp 0   TMA producer (kStages-deep smem ring, mbarrier full/empty)
//   warp 1   tcgen05.mma issuer (one elected lane), commits to empty / tmem_full
//   warp 2   TMEM allocator (2 accumulator stages -> epilogue overlaps next tile's MMA)
//   warps 4-7 epilogue: tcgen05.ld -> bias/ReLU/mask or softmax -> bf16 -> swizzled smem
//             -> TMA store (clips the ragged last tile of every utterance)
//
// Split-bf16 ("bf16x2") mode: activations and weights carry a second bf16 plane holding
// x - bf16(x); the K loop then issues hi*hi + hi*lo + lo*hi (3 MMAs per chunk), which
// restores ~16 mantissa bits and lets the tensor-core path meet the fp32 parity
// tolerance (logits <= 1e-3 rel) that plain bf16 cannot.
#include "conv_umma.h"

namespace sl {

using namespace ptx;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements per smem row = 128 B = one swizzle span
constexpr int UMMA_K = 16;
// warps 0-3: TMA producer, MMA issuer, TMEM allocator, idle; then the epilogue warps.  A lone warp
// per scheduler issues one of its (mostly dependent) instructions every ~2.6 cycles (measured), so
// the bias / ReLU / mask / pack epilogue runs on TWO warps per TMEM lane quadrant, which share the
// 64-column chunks of a tile alternately; the softmax epilogue (one chunk per tile) keeps one.
template <int EPI>
struct Threads {
  static constexpr int kEpiSets = EPI == EPI_SOFTMAX ? 1 : 2;
  static constexpr int kEpiThreads = 128 * kEpiSets;
  static constexpr int kThreads = 128 + kEpiThreads;
};
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int STAGING_BYTES = BLOCK_M * 128;  // 128 rows x 128 B
constexpr int WARP_STAGING_BYTES = 32 * 128;  // one epilogue warp's 32-row slab

// CTAS = 2: CTA pair (cta_group::2).  The pair computes a 256-frame x BN-filter tile; each CTA stages
// its own 128 frames of A and BN / 2 filters of B, and the leader's MMAs (M = 256) read both halves.
template <int BN, int CTAS>
struct Cfg {
  static constexpr int BN_LOCAL = BN / CTAS;  // filters of the B tile this CTA stages
  static constexpr int B_BYTES = BN_LOCAL * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int kStages = (BN >= 256 && CTAS == 1) ? 4 : 6;
  static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : 2 * BN;
  // stages | 2 staging buffers | bias | barriers
  static constexpr int SMEM_BYTES =
      kStages * STAGE_BYTES + 2 * STAGING_BYTES + BN * 4 + 512 + 1024 /*align slack*/;
  // halo mode re-carves the stage region: two halo buffers of HALO_BYTES, then B-only stages
  static constexpr int HALO_BYTES = 32768;  // up to 256 frames x 128 B
  static constexpr int kStagesB = (kStages * STAGE_BYTES - 2 * HALO_BYTES) / B_BYTES > 8
                                      ? 8
                                      : (kStages * STAGE_BYTES - 2 * HALO_BYTES) / B_BYTES;
};

// Work item of the persistent loop: an output tile, or one of the tail_split narrow slices of a
// tile of the last partial wave.
struct WorkItem {
  int b, t0, n0, width, chunk_begin, chunk_end;  // [chunk_begin, chunk_end): 64-channel chunks of the contraction
  int tail;  // tail K split: index of the tile within the partial wave (its scratch slot), else -1
};
// `rank`: this CTA's rank in its pair (0 without pairs); a pair works on the frame tiles
// 2 * unit and 2 * unit + 1 of one filter tile.  The odd tile out (if any) gets b = p.B: its loads
// are zero-filled and its stores clipped by the tensor maps.
template <int BN, int CTAS>
__device__ __forceinline__ WorkItem decode_item(const ConvGemmParams& p, int item, int rank) {
  int tile = item, sub = 0, width = BN, tail = -1;
  int chunk_begin = 0, chunk_end = p.chunks;
  if (p.ksplit > 1) {
    // Split K over CHANNEL chunks (all taps each), not over tap ranges: a tap-range item still needs
    // the activation rows of (almost) the whole tile for every channel, so ksplit tap-range items read
    // the A operand ksplit times (ncu on the big_conv_1 input gradient: 900 MB of DRAM reads for a 164 MB
    // tensor); channel ranges partition both operands.  Split-major order: CTAs running together share a
    // chunk range (the same weight slices in L2).
    const int total = p.m_units * p.n_tiles;
    const int split = item / total;
    tile = item - split * total;
    const int per = (p.chunks + p.ksplit - 1) / p.ksplit;
    chunk_begin = split * per < p.chunks ? split * per : p.chunks;
    chunk_end = chunk_begin + per < p.chunks ? chunk_begin + per : p.chunks;
  } else if (item >= p.full_tiles && p.tail_ksplit > 1) {
    const int r = item - p.full_tiles;
    tail = r / p.tail_ksplit;
    tile = p.full_tiles + tail;
    chunk_begin = (r - tail * p.tail_ksplit) * p.tail_per;
    chunk_end = chunk_begin + p.tail_per < p.chunks ? chunk_begin + p.tail_per : p.chunks;
  } else if (item >= p.full_tiles) {
    const int r = item - p.full_tiles;
    tile = p.full_tiles + r / p.tail_split;
    sub = r - (r / p.tail_split) * p.tail_split;
    width = BN / p.tail_split;
  }
  if (p.reverse_order) tile = p.m_units * p.n_tiles - 1 - tile;  // (whole tiles only, see conv_gemm_launch)
  // filter tile fastest: the n_tiles CTAs that share an activation tile run together, so the
  // A operand comes from HBM once and from L2 afterwards (the weights are L2 resident anyway;
  // with the filter tile slowest, big_conv_2 re-read its 164 MB input 8 times from HBM)
  const int n_tile = tile % p.n_tiles;
  const int rem = (tile / p.n_tiles) * CTAS + rank;
  WorkItem w;
  w.b = rem / p.m_tiles_per_utt;
  w.t0 = (rem - w.b * p.m_tiles_per_utt) * BLOCK_M;
  if (CTAS > 1 && w.b >= p.B) {
    w.b = p.B;
    w.t0 = 0;
  }
  w.n0 = n_tile * BN + sub * width;
  w.width = width;
  w.chunk_begin = chunk_begin;
  w.chunk_end = chunk_end;
  w.tail = tail;
  return w;
}

// BMN: the B operand is MN-major (input gradient: B[n = ci][k = co] read straight from the
// forward weight layout (k, co, ci), ci contiguous) instead of K-major.
template <int CTAS>
__device__ __forceinline__ void tma3(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  if constexpr (CTAS == 2)
    tma_load_3d_pair(m, mapa_u32(smem_u32(bar), 0), dst, c0, c1, c2);  // counted on the leader's barrier
  else
    tma_load_3d(m, bar, dst, c0, c1, c2);
}
template <int CTAS>
__device__ __forceinline__ void tma4(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                     int c3) {
  if constexpr (CTAS == 2)
    tma_load_4d_pair(m, mapa_u32(smem_u32(bar), 0), dst, c0, c1, c2, c3);
  else
    tma_load_4d(m, bar, dst, c0, c1, c2, c3);
}
// MMA-retired signal: to this CTA's barrier, or to the barrier at the same offset in both CTAs of a pair
template <int CTAS>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if constexpr (CTAS == 2)
    umma_commit_pair(bar);
  else
    umma_commit(bar);
}
// epilogue warp -> MMA issuer (always in the leader CTA): accumulator stage drained
template <int CTAS>
__device__ __forceinline__ void arrive_leader(uint64_t* bar) {
  if constexpr (CTAS == 2)
    mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
  else
    mbar_arrive(bar);
}

// tail K split: counters shared by the work items of one tile
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// TAILK: built with the tail-K-split epilogue paths (ConvGemmParams::tail_ksplit); a separate instantiation so
// that the default kernel does not carry their registers
template <int BN, int EPI, bool BMN, int CTAS, bool TAILK>
__global__ void __launch_bounds__(Threads<EPI>::kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BN, CTAS>;
  constexpr int kEpiSets = Threads<EPI>::kEpiSets;
  constexpr int kEpiThreads = Threads<EPI>::kEpiThreads;
  const int rank = CTAS == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cta = CTAS == 2 ? static_cast<int>(blockIdx.x) / CTAS : static_cast<int>(blockIdx.x);  // work-item lane
  const int n_ctas = static_cast<int>(gridDim.x) / CTAS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* staging = smem + C::kStages * C::STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(staging + 2 * STAGING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + BN);
  // barrier block: [0,8) full, [8,16) empty (ring of A+B stages, or of B-only stages in halo
  // mode), [16,18) halo full, [18,20) halo empty, [20,22) tmem full, [22,24) tmem empty
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 8;
  uint64_t* halo_full = bars + 16;
  uint64_t* halo_empty = bars + 18;
  uint64_t* tmem_full = bars + 20;
  uint64_t* tmem_empty = bars + 22;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 24);
  volatile uint32_t* tail_role_s = reinterpret_cast<volatile uint32_t*>(bars + 26);  // [2], by accumulator stage
  static_assert(C::kStages <= 8 && C::kStagesB <= 8 && C::kStagesB >= 2, "barrier block layout");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto stamp = [&](int slot) {  // measurement aid (SL_TIMELINE=1, api.cu)
    if (p.timeline != nullptr && slot < 32) p.timeline[blockIdx.x * 32 + slot] = clock64();
  };
  if (threadIdx.x == 0) stamp(0);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    if (p.halo) prefetch_tmap(&p.tmAhalo);
    if (p.tail_split > 1) prefetch_tmap(&p.tmBtail);
    if (TAILK && p.tail_ksplit > 1) prefetch_tmap(&p.tmScratch);
    if (EPI != EPI_SOFTMAX) prefetch_tmap(&p.tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&halo_full[a], 1);
      mbar_init(&halo_empty[a], 1);
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4 * kEpiSets * CTAS);  // every epilogue warp of every CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTAS == 2) {
      tmem_alloc_pair(tmem_ptr_s, C::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_s, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tcgen05_fence_before();
  if constexpr (CTAS == 2)
    cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  else
    __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (threadIdx.x == 0) stamp(1);
  pdl_wait();  // setup above overlapped the previous kernel's tail; its outputs are visible from here
  if (threadIdx.x == 0) stamp(2);

  const int total_tiles = p.m_units * p.n_tiles;
  const int num_tiles = p.ksplit > 1 ? total_tiles * p.ksplit
                                     : p.full_tiles + (total_tiles - p.full_tiles) *
                                                          (p.tail_ksplit > 1 ? p.tail_ksplit : p.tail_split);  // work items

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int issued = 0;
      int hbuf = 0;
      uint32_t hphase = 0;
      for (int tile = cta; tile < num_tiles; tile += n_ctas) {
        const WorkItem w = decode_item<BN, CTAS>(p, tile, rank);
        const int b = w.b, t0 = w.t0, n0 = w.n0;
        const CUtensorMap* tmB = w.width == BN ? &p.tmB : &p.tmBtail;
        const int n_loc = n0 + rank * (w.width / CTAS);  // first filter of the B half this CTA stages
        const bool expects = CTAS == 1 || rank == 0;     // a pair counts all its bytes on the leader's barrier
        const uint32_t stage_tx = A_BYTES + (w.width / CTAS) * (BLOCK_K * 2);
        if (p.halo) {
          // chunk-major: one halo tile per (chunk, term), then one B tile per tap
          uint8_t* b_ring = smem + 2 * C::HALO_BYTES;
          const uint32_t b_tx = (w.width / CTAS) * (BLOCK_K * 2);
          for (int chunk = w.chunk_begin; chunk < w.chunk_end; ++chunk) {
            for (int term = 0; term < p.terms; ++term) {
              const int a_c = chunk * BLOCK_K + (term == 2 ? p.a_lo_off : 0);
              const int b_c = chunk * BLOCK_K + (term == 1 ? p.b_lo_off : 0);
              mbar_wait(&halo_empty[hbuf], hphase ^ 1);
              if (expects) mbar_expect_tx(&halo_full[hbuf], CTAS * p.halo_rows * 128);
              tma4<CTAS>(&p.tmAhalo, &halo_full[hbuf], smem + hbuf * C::HALO_BYTES, a_c, 0, t0 - p.pad_l, b);
              if (++hbuf == 2) {
                hbuf = 0;
                hphase ^= 1;
              }
              for (int tap = 0; tap < p.taps; ++tap) {
                const int wtap = p.w_tap0 + tap * p.w_tap_step;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* b_s = b_ring + stage * C::B_BYTES;
                if (expects) mbar_expect_tx(&full_bar[stage], CTAS * b_tx);
                if (BMN)
                  tma4<CTAS>(tmB, &full_bar[stage], b_s, 0, chunk * BLOCK_K,
                             (n_loc + (term == 1 ? p.b_lo_off : 0)) >> 6, wtap);
                else
                  tma3<CTAS>(tmB, &full_bar[stage], b_s, b_c, n_loc, wtap);
                if (++stage == C::kStagesB) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
          continue;
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int jp = tap - p.pad_l;  // signed frame offset of this tap
          int q, par;
          if (p.stride == 1) {
            q = jp;
            par = 0;
          } else {  // floor division / non-negative modulo by the stride
            q = jp >= 0 ? jp / p.stride : -((-jp + p.stride - 1) / p.stride);
            par = jp - q * p.stride;
          }
          const int wtap = p.w_tap0 + tap * p.w_tap_step;
          for (int chunk = w.chunk_begin; chunk < w.chunk_end; ++chunk) {
            for (int term = 0; term < p.terms; ++term) {
              const int a_c = chunk * BLOCK_K + (term == 2 ? p.a_lo_off : 0);
              const int b_c = chunk * BLOCK_K + (term == 1 ? p.b_lo_off : 0);
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* a_s = smem + stage * C::STAGE_BYTES;
              uint8_t* b_s = a_s + A_BYTES;
              if (CTAS == 1 && p.dbg_mode == 1 && issued >= C::kStages) {  // measurement aid: MMA on stale tiles
                mbar_arrive(&full_bar[stage]);
                if (++stage == C::kStages) {
                  stage = 0;
                  phase ^= 1;
                }
                continue;
              }
              ++issued;
              if (CTAS == 1 && p.dbg_mode >= 2 && issued > C::kStages) {
                // measurement aids: 2 = keep only the A loads (what a shared-B / 2-CTA scheme would
                // save), 3 = keep only the B loads (what reusing the A halo across taps would save)
                if (p.dbg_mode == 2) {
                  mbar_expect_tx(&full_bar[stage], A_BYTES);
                  tma_load_4d(&p.tmA, &full_bar[stage], a_s, a_c, par, t0 + q, b);
                } else {
                  mbar_expect_tx(&full_bar[stage], stage_tx - A_BYTES);
                  if (BMN)
                    tma_load_4d(tmB, &full_bar[stage], b_s, 0, chunk * BLOCK_K,
                                (n0 + (term == 1 ? p.b_lo_off : 0)) >> 6, wtap);
                  else
                    tma_load_3d(tmB, &full_bar[stage], b_s, b_c, n0, wtap);
                }
                if (++stage == C::kStages) {
                  stage = 0;
                  phase ^= 1;
                }
                continue;
              }
              if (expects) mbar_expect_tx(&full_bar[stage], CTAS * stage_tx);
              tma4<CTAS>(&p.tmA, &full_bar[stage], a_s, a_c, par, t0 + q, b);
              if (BMN) {
                // 64 contraction rows (co) x 64 output channels (ci) per box
                const int n_c = n_loc + (term == 1 ? p.b_lo_off : 0);
                if (p.b_grouped || CTAS == 2) {
                  // one box {64 ci, 64 co rows, BN_LOCAL/64 channel groups}: lands as [group][row][64]
                  tma4<CTAS>(tmB, &full_bar[stage], b_s, 0, chunk * BLOCK_K, n_c >> 6, wtap);
                } else {
#pragma unroll
                  for (int i = 0; i < BN / 64; ++i)
                    if (i * 64 < w.width)
                      tma_load_3d(&p.tmB, &full_bar[stage], b_s + i * (BLOCK_K * 128), n_c + 64 * i,
                                  chunk * BLOCK_K, wtap);
                }
              } else {
                tma3<CTAS>(tmB, &full_bar[stage], b_s, b_c, n_loc, wtap);
              }
              if (++stage == C::kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 && (CTAS == 1 || rank == 0)) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    int hbuf = 0;
    uint32_t hphase = 0;
    for (int tile = cta; tile < num_tiles; tile += n_ctas, ++it) {
      const WorkItem w = decode_item<BN, CTAS>(p, tile, rank);
      const int ksteps = p.taps * (w.chunk_end - w.chunk_begin) * p.terms;
      const uint32_t idesc = make_idesc_16(BLOCK_M * CTAS, w.width, 0, BMN ? 1 : 0, p.fp16);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
      if (p.halo) {
        const uint32_t b_ring = smem_u32(smem + 2 * C::HALO_BYTES);
        const int ntaps = p.taps;
        const int n_ct = (w.chunk_end - w.chunk_begin) * p.terms;
        uint32_t first = 1;
        for (int ct = 0; ct < n_ct; ++ct) {
          mbar_wait(&halo_full[hbuf], hphase);
          const uint32_t halo_addr = smem_u32(smem + hbuf * C::HALO_BYTES);
          for (int ti = 0; ti < ntaps; ++ti) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            if (elect_one()) {  // (a single-taker branch ptxas can see: operands stay in uniform registers)
              // tap j reads halo rows [j, j + 128): start address advanced by j rows of 128 B
              const uint32_t a_addr = halo_addr + static_cast<uint32_t>(ti) * 128u;
              const uint32_t b_addr = b_ring + stage * C::B_BYTES;
              const uint64_t db0 = BMN ? make_smem_desc_sw128(b_addr, BLOCK_K * 128, 1024)
                                       : make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                const uint32_t aa = a_addr + k * UMMA_K * 2;
                const uint64_t da = make_smem_desc_sw128_off(aa, 16, 1024, p.halo_base_mode ? (aa >> 7) & 7u : 0u);
                const uint64_t db = db0 + static_cast<uint64_t>(k * (BMN ? (2048 >> 4) : ((UMMA_K * 2) >> 4)));
                if constexpr (CTAS == 2)
                  umma_bf16_pair(tmem_d, da, db, idesc, first ? 0u : 1u);
                else
                  umma_bf16(tmem_d, da, db, idesc, first ? 0u : 1u);
                first = 0;
              }
              commit<CTAS>(&empty_bar[stage]);
              if (ti == ntaps - 1) commit<CTAS>(&halo_empty[hbuf]);
              if (ti == ntaps - 1 && ct == n_ct - 1) commit<CTAS>(&tmem_full[as]);
            }
            __syncwarp();
            first = 0;
            if (++stage == C::kStagesB) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++hbuf == 2) {
            hbuf = 0;
            hphase ^= 1;
          }
        }
        if (ksteps == 0 && lane == 0) {
          mbar_arrive(&tmem_full[as]);
          if (CTAS == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_full[as]), 1));
        }
        continue;
      }
      if constexpr (CTAS == 1) {
        // The issue loop is bound by this one thread's instruction stream, not by the tensor pipe (§7.1 of
        // four extra clock reads per MMA cost +85 cycles each, and only two MMAs are ever in
        // flight), so it is kept as short as it gets: the leader is chosen with elect.sync — a branch ptxas
        // knows to have a single taker, so the operands can live in uniform registers — and the four
        // descriptors of a stage differ only in their 14-bit address field (one add each).
        const uint32_t ring = smem_u32(smem);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (lane == 0 && ks == 0) stamp(it == 0 ? 3 : 25 + it);
          if (elect_one()) {
            const uint32_t a_addr = ring + stage * C::STAGE_BYTES;
            const uint64_t da0 = make_smem_desc_sw128(a_addr, 16, 1024);
            const uint64_t db0 = BMN ? make_smem_desc_sw128(a_addr + A_BYTES, BLOCK_K * 128, 1024)
                                     : make_smem_desc_sw128(a_addr + A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // K-major: 16 contraction elements = 32 B further along the swizzled row; MN-major B: two 8-row
              // swizzle atoms (2048 B) per 16 contraction rows
              const uint64_t da = da0 + static_cast<uint64_t>(k * ((UMMA_K * 2) >> 4));
              const uint64_t db = db0 + static_cast<uint64_t>(k * (BMN ? (2048 >> 4) : ((UMMA_K * 2) >> 4)));
              umma_bf16(tmem_d, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (ks == ksteps - 1) umma_commit(&tmem_full[as]);
          }
          __syncwarp();
          if (lane == 0 && ks == ksteps - 1 && it < 6) stamp(4 + it);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      } else {
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {  // (as in the single-CTA loop: single-taker branch, descriptors one add apart)
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t da0 = make_smem_desc_sw128(a_addr, 16, 1024);
          // MN-major B: 16 contraction rows = two 8-row swizzle atoms (2048 B); 64-channel
          // groups are BLOCK_K * 128 B apart (leading byte offset)
          const uint64_t db0 = BMN ? make_smem_desc_sw128(a_addr + A_BYTES, BLOCK_K * 128, 1024)
                                   : make_smem_desc_sw128(a_addr + A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = da0 + static_cast<uint64_t>(k * ((UMMA_K * 2) >> 4));
            const uint64_t db = db0 + static_cast<uint64_t>(k * (BMN ? (2048 >> 4) : ((UMMA_K * 2) >> 4)));
            umma_bf16_pair(tmem_d, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          commit<CTAS>(&empty_bar[stage]);  // frees the smem slot (of both CTAs) when these MMAs retire
          if (ks == ksteps - 1) commit<CTAS>(&tmem_full[as]);
        }
        __syncwarp();
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      }
      if (ksteps == 0 && lane == 0) {  // empty tap range (never planned)
        mbar_arrive(&tmem_full[as]);
        if (CTAS == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_full[as]), 1));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp & 3;          // TMEM lane quadrant this warp may read
    const int set = (warp - 4) >> 2;  // which of the kEpiSets warps of the quadrant: takes chunks set, set + 2, ...
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 128;
    int it = 0;
    int staged_n0 = -1;
    // one 32-row x 128-byte staging slab per warp: by the time a warp has prepared its next chunk,
    // the TMA store of the previous one has long finished reading the slab
    uint8_t* const sbuf = staging + (warp - 4) * WARP_STAGING_BYTES;
    for (int tile = cta; tile < num_tiles; tile += n_ctas, ++it) {
      const WorkItem w = decode_item<BN, CTAS>(p, tile, rank);
      const int b = w.b, t0 = w.t0, n0 = w.n0;
      const int n_chunks = w.width / 64;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int t = t0 + row;
      const bool row_valid = t < p.T_out && b < p.B;

      // stage the bias slice of this tile; consecutive tiles of a CTA mostly share their filter
      // range, so the (two-barrier) refill only runs when it changes
      if (n0 != staged_n0) {
        named_bar_sync(2, kEpiThreads);  // every warp is done reading the previous slice
        for (int i = et; i < BN; i += kEpiThreads)
          bias_s[i] = (p.bias != nullptr && (n0 + i) < p.n_valid) ? p.bias[n0 + i] : 0.f;
        named_bar_sync(2, kEpiThreads);
        staged_n0 = n0;
      }
      // ReLU bitmask of the whole row slice, fetched before the accumulator is ready
      uint2 mask_in[BN / 64];
      if (EPI == EPI_PACKED && p.mask_bits_in != nullptr && row_valid) {
        const uint2* mp = reinterpret_cast<const uint2*>(
            p.mask_bits_in +
            (static_cast<size_t>(b) * p.mask_T + (t * p.out_t_scale + p.out_t_off)) * p.mask_row_bytes + (n0 >> 3));
#pragma unroll
        for (int c = 0; c < BN / 64; ++c)
          if (c < n_chunks && (c % kEpiSets) == set) mask_in[c] = __ldg(mp + c);
      }

      // tail K split: 1 = this item adds its partial sums to the tile's scratch slot; 2 = it drew the last
      // ticket: it waits for the others' sums, folds them into its accumulator and runs the epilogue
      int tail_role = 0;
      if constexpr (TAILK) {
        if (w.tail >= 0) {
          if (et == 0) {
            const unsigned splits = static_cast<unsigned>(p.tail_ksplit);
            const unsigned ticket = atomicInc(p.tail_tickets + w.tail, splits - 1);  // wraps to 0 after the last
            if (ticket == splits - 1) {
              while (ld_acquire_gpu_u32(p.tail_done + w.tail) != splits - 1) __nanosleep(64);
              p.tail_done[w.tail] = 0;  // nobody else touches it before the next launch
              fence_proxy_async_all();  // the sums were written through the async proxy
              tail_role_s[as] = 2;
            } else {
              tail_role_s[as] = 1;
            }
          }
          named_bar_sync(2, kEpiThreads);
          tail_role = static_cast<int>(tail_role_s[as]);
        }
      }

      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      if (warp == 4 && lane == 0 && it < 6) stamp(10 + it);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                             static_cast<uint32_t>(as * BN);

      if (EPI == EPI_F32) {
        // split-K partial sums: fp32 accumulator -> per-warp swizzled slab -> TMA reduce-add
        const bool has_data = w.chunk_end > w.chunk_begin;
#pragma unroll 1
        for (int c = set; c < (has_data ? w.width / 32 : 0); c += kEpiSets) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (lane == 0) tma_wait_group_read<0>();
          __syncwarp();
          uint8_t* rowp = sbuf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int phys = j ^ (lane & 7);
            *reinterpret_cast<uint4*>(rowp + phys * 16) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_3d(&p.tmY, sbuf, n0 + c * 32, t0 + ew * 32, b);
            tma_commit_group();
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader<CTAS>(&tmem_empty[as]);
      } else if (TAILK && tail_role == 1) {
        // fp32 accumulator -> per-warp swizzled slab -> TMA reduce-add into the tile's scratch slot
#pragma unroll 1
        for (int c = set; c < BN / 32; c += kEpiSets) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (lane == 0) tma_wait_group_read<0>();
          __syncwarp();
          uint8_t* rowp = sbuf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int phys = j ^ (lane & 7);
            *reinterpret_cast<uint4*>(rowp + phys * 16) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_3d(&p.tmScratch, sbuf, c * 32, ew * 32, w.tail);
            tma_commit_group();
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          arrive_leader<CTAS>(&tmem_empty[as]);
          tma_wait_group<0>();  // the sums have landed, not just left the slab
          fence_proxy_async_all();
          __threadfence();
        }
        named_bar_sync(2, kEpiThreads);
        if (et == 0) {
          __threadfence();
          red_release_gpu_add_u32(p.tail_done + w.tail, 1u);
        }
      } else if (EPI == EPI_PACKED) {
        if (set >= n_chunks) {  // narrow tail tile: nothing for this warp, but the MMA warp counts every arrival
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader<CTAS>(&tmem_empty[as]);
        }
#pragma unroll 1
        for (int c = set; c < n_chunks; c += kEpiSets) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + c * 64, r0);
          tmem_ld_32x32(taddr + c * 64 + 32, r1);
          tmem_ld_wait();
          if (c + kEpiSets >= n_chunks) {
            // this warp's share of the accumulator is drained: hand the TMEM stage back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader<CTAS>(&tmem_empty[as]);
          }
          float v[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = __uint_as_float(r0[i]);
            v[32 + i] = __uint_as_float(r1[i]);
          }
          if (TAILK && tail_role == 2) {
            // the other items' partial sums: read coalesced (and replaced by zeros, so that the slot is clean
            // for the next launch), transposed through the warp's slab so that every lane gets its own row
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float* base = p.tail_scratch + (static_cast<size_t>(w.tail) * BLOCK_M + ew * 32) * BN + c * 64 + h * 32;
              if (lane == 0) tma_wait_group_read<0>();  // the store that last read sbuf is done
              __syncwarp();
              float4 q[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4* g = reinterpret_cast<float4*>(base + (j * 4 + (lane >> 3)) * BN) + (lane & 7);
                q[j] = __ldcg(g);
                __stcg(g, zero4);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int rr = j * 4 + (lane >> 3);
                *reinterpret_cast<float4*>(sbuf + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4)) = q[j];
              }
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = *reinterpret_cast<const float4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4));
                v[h * 32 + 4 * j] += t.x;
                v[h * 32 + 4 * j + 1] += t.y;
                v[h * 32 + 4 * j + 2] += t.z;
                v[h * 32 + 4 * j + 3] += t.w;
              }
              __syncwarp();
            }
          }
          if (p.bias != nullptr) {  // (uniform) the input-gradient GEMMs have none
            const float4* bs = reinterpret_cast<const float4*>(bias_s + c * 64);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float4 q = bs[i];
              v[4 * i] += q.x;
              v[4 * i + 1] += q.y;
              v[4 * i + 2] += q.z;
              v[4 * i + 3] += q.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (p.mask_bits_in != nullptr && row_valid) {
            // ReLU backward: keep the gradient only where the forward output was > 0
            uint2 mb = mask_in[0];
#pragma unroll
            for (int cc = 1; cc < BN / 64; ++cc)
              if (cc == c) mb = mask_in[cc];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (!((mb.x >> i) & 1u)) v[i] = 0.f;
              if (!((mb.y >> i) & 1u)) v[32 + i] = 0.f;
            }
          }
          if (p.out_scale != 1.0f) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] *= p.out_scale;
          }
          if (p.mask_bits_out != nullptr && row_valid) {
            uint2 mb = make_uint2(0u, 0u);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              mb.x |= (v[i] > 0.f ? 1u : 0u) << i;
              mb.y |= (v[32 + i] > 0.f ? 1u : 0u) << i;
            }
            *reinterpret_cast<uint2*>(p.mask_bits_out +
                                      (static_cast<size_t>(b) * p.T_out + t) * p.mask_row_bytes +
                                      ((n0 + c * 64) >> 3)) = mb;
          }
          // Each epilogue warp owns a 32-row slab: it stages and TMA-stores it on its own, so the
          // warps never wait for each other (bulk async-groups are per thread).
          auto stage_and_store = [&](const uint32_t (&packed)[32], int col0) {
            if (lane == 0) tma_wait_group_read<0>();  // the store that last read sbuf is done
            __syncwarp();
            uint8_t* rowp = sbuf + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int phys = j ^ (lane & 7);  // SWIZZLE_128B: 16-byte chunk index ^= row % 8
              *reinterpret_cast<uint4*>(rowp + phys * 16) =
                  make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&p.tmY, sbuf, col0, t0 + ew * 32, b);
              tma_commit_group();
            }
          };
          {
            uint32_t packed[32];
            if (p.fp16) {  // (uniform)
#pragma unroll
              for (int i = 0; i < 32; ++i) packed[i] = pack_fp16x2(v[2 * i], v[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) packed[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            }
            stage_and_store(packed, n0 + c * 64);
          }
          if (p.y_planes == 2) {  // (uniform) split-bf16 mode only: the residual plane x - bf16(x)
            uint32_t packed[32];
#pragma unroll
            for (int i = 0; i < 32; ++i)
              packed[i] = pack_bf16x2(v[2 * i] - bf16_round(v[2 * i]), v[2 * i + 1] - bf16_round(v[2 * i + 1]));
            stage_and_store(packed, n0 + c * 64 + p.y_lo_off);
          }
        }
      } else {
        // EPI_SOFTMAX: BN == 64, one row of V logits per thread (net.py:328-330)
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(taddr, r0);
        tmem_ld_32x32(taddr + 32, r1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader<CTAS>(&tmem_empty[as]);
        if (row_valid) {
          float z[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            z[i] = __uint_as_float(r0[i]) + bias_s[i];
            z[32 + i] = __uint_as_float(r1[i]) + bias_s[32 + i];
          }
          const int V = p.V;
          float m = -INFINITY;
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i < V) m = fmaxf(m, z[i]);
          float e[64];
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            e[i] = i < V ? expf(z[i] - m) : 0.f;
            sum += e[i];
          }
          const float inv = 1.f / sum;
          float sum_eps = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            e[i] *= inv;
            if (i < V) sum_eps += e[i] + 1e-8f;
          }
          const size_t ro = static_cast<size_t>(b) * p.T_out + t;
          float* pr = p.probs + ro * V;
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i < V) pr[i] = e[i];
          if (p.logits != nullptr) {
            float* lg = p.logits + ro * V;
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (i < V) lg[i] = z[i];
          }
          if (p.logp != nullptr) {
            // K.ctc_batch_cost feeds log(p + 1e-8) to tf.nn.ctc_loss, which applies its own
            // log-softmax: logp = log(p+eps) - log(sum_v (p_v+eps))      
            const float lse = logf(sum_eps);
            float* lp = p.logp + ro * 64;
#pragma unroll
            for (int i = 0; i < 64; ++i) lp[i] = i < V ? logf(e[i] + 1e-8f) - lse : -INFINITY;
          }
        }
      }
      if (warp == 4 && lane == 0 && it < 6) stamp(16 + it);
    }
    if (EPI != EPI_SOFTMAX && lane == 0) tma_wait_group<0>();
    if (warp == 4 && lane == 0) stamp(22);
  }

  tcgen05_fence_before();
  if constexpr (CTAS == 2)
    cluster_sync_all();  // neither CTA may exit while the other can still signal its barriers
  else
    __syncthreads();
  if (threadIdx.x == 0) stamp(23);
  if (warp == 2) {
    tcgen05_fence_after();
    if constexpr (CTAS == 2)
      tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else
      tmem_dealloc(tmem_base, C::TMEM_COLS);
    if (lane == 0) stamp(24);
  }
}

template <int BN, int EPI, bool BMN, int CTAS, bool TAILK = false>
int launch(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<BN, CTAS>;
  // the opt-in shared-memory size is a per-device function attribute
  static unsigned long long configured = 0;  // bit d: set for device d (benign race: idempotent)
  static int max_clusters[64] = {0};         // CTA pairs that can be co-resident on device d
  int dev = 0;
  SL_CUDA(cudaGetDevice(&dev));
  if (!((configured >> (dev & 63)) & 1ull)) {
    SL_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, EPI, BMN, CTAS, TAILK>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured |= 1ull << (dev & 63);
  }
  const int num_tiles = p.m_units * p.n_tiles * (p.ksplit > 1 ? p.ksplit : 1);
  if (CTAS == 2) {
    if (p.tail_split != 1 || p.tail_ksplit > 1 || p.full_tiles != p.m_units * p.n_tiles || p.dbg_mode != 0) {
      set_error("conv_gemm: the CTA-pair kernel takes whole tiles only");
      return 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(Threads<EPI>::kThreads);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int& mc = max_clusters[dev & 63];
    if (mc == 0) {
      cfg.gridDim = dim3(num_sms & ~1);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<BN, EPI, BMN, CTAS, TAILK>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = num_sms / 2;
      }
      mc = n;
    }
    const int pairs = num_tiles < mc ? num_tiles : mc;
    cfg.gridDim = dim3(2 * pairs);
    SL_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, EPI, BMN, CTAS, TAILK>, p));
    return 0;
  }
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  if (p.ksplit <= 1 && (p.full_tiles + (num_tiles - p.full_tiles) * (p.tail_ksplit > 1 ? p.tail_ksplit : p.tail_split) < grid ||
                        p.tail_split < 1 || (p.tail_ksplit > 1) != TAILK || (TAILK && p.tail_split != 1))) {
    set_error("conv_gemm: inconsistent tail split");
    return 1;
  }
  SL_CUDA(launch_pdl(PDL_CONV, conv_gemm_kernel<BN, EPI, BMN, CTAS, TAILK>, dim3(grid), dim3(Threads<EPI>::kThreads), C::SMEM_BYTES, stream,
                     p));
  return 0;
}

}  // namespace

int conv_gemm_launch(const ConvGemmParams& p, int block_n, int epi, bool b_mn_major, int num_sms,
                     cudaStream_t stream) {
  if (p.ctas == 2) {
    SL_REQUIRE(block_n == 256 && epi != EPI_SOFTMAX, "the CTA-pair kernel is built for 256-filter tiles");
    SL_REQUIRE(!b_mn_major || p.b_grouped, "the CTA-pair kernel needs the grouped weight map");
    if (epi == EPI_F32) {
      SL_REQUIRE(b_mn_major, "split-K epilogue is built for the dgrad tiles");
      return launch<256, EPI_F32, true, 2>(p, num_sms, stream);
    }
    return b_mn_major ? launch<256, EPI_PACKED, true, 2>(p, num_sms, stream)
                      : launch<256, EPI_PACKED, false, 2>(p, num_sms, stream);
  }
  SL_REQUIRE(p.ctas == 1, "ctas must be 1 or 2");
  SL_REQUIRE(!p.reverse_order || (p.ksplit <= 1 && p.tail_split == 1 && p.tail_ksplit <= 1 &&
                                  p.full_tiles == p.m_units * p.n_tiles),
             "reverse tile order is for whole tiles");
  if (epi == EPI_SOFTMAX) {
    SL_REQUIRE(block_n == 64 && !b_mn_major, "softmax epilogue needs a 64-wide K-major tile");
    return launch<64, EPI_SOFTMAX, false, 1>(p, num_sms, stream);
  }
  if (epi == EPI_F32) {
    SL_REQUIRE(b_mn_major && (block_n == 128 || block_n == 256), "split-K epilogue is built for the dgrad tiles");
    return block_n == 256 ? launch<256, EPI_F32, true, 1>(p, num_sms, stream)
                          : launch<128, EPI_F32, true, 1>(p, num_sms, stream);
  }
  switch (block_n) {
    case 64:
      return b_mn_major ? launch<64, EPI_PACKED, true, 1>(p, num_sms, stream)
                        : launch<64, EPI_PACKED, false, 1>(p, num_sms, stream);
    case 128:
      return b_mn_major ? launch<128, EPI_PACKED, true, 1>(p, num_sms, stream)
                        : launch<128, EPI_PACKED, false, 1>(p, num_sms, stream);
    case 256:
      if (p.tail_ksplit > 1)
        return b_mn_major ? launch<256, EPI_PACKED, true, 1, true>(p, num_sms, stream)
                          : launch<256, EPI_PACKED, false, 1, true>(p, num_sms, stream);
      return b_mn_major ? launch<256, EPI_PACKED, true, 1>(p, num_sms, stream)
                        : launch<256, EPI_PACKED, false, 1>(p, num_sms, stream);
    default:
      set_error("conv_gemm: unsupported block_n");
      return 1;
  }
}

}  // namespace sl
