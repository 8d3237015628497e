// wgrad_umma.cu — Conv1D weight gradient on tcgen05 (TF autodiff of net.py:304-305):
//
//   dW[j, co, ci] = sum_{b,t} dY[b, t, co] * X[b, t*s + j - pad_l, ci]
//
// The contraction runs over time, which is the *slow* axis of both channels-last
// tensors, so both UMMA operands are MN-major: a TMA box of 64 frames x 64 channels
// lands as 64 rows of 128 B (SWIZZLE_128B) and is consumed directly with the
// transposing smem descriptors (a_major = b_major = MN).  As in the forward kernel the
// tap shift j and the SAME padding are signed TMA coordinates + OOB zero fill.
//
//   D[128 co x BN ci] (TMEM fp32) += dY_tile^T[128 co x 64 t] * X_tile[64 t x BN ci]
//
// Work unit = (tap, 128-filter tile, BN-channel tile, K split); the K range of a unit
// is a contiguous run of (utterance, 64-frame chunk) pairs.  The epilogue stages the fp32
// accumulator through swizzled smem, 32 columns at a time, and hands it to TMA: a plain
// tensor store, or cp.reduce.async.bulk.tensor (.add) when several K splits / layers of
// accumulation meet in the same dW tile (dW is zeroed by the caller in that case).
//
// Bias gradient, fused: db[co] = sum_{b,t} dY[b,t,co] is the column sum of the very dY tiles
// the A operand streams through smem, so the otherwise idle warp 3 adds up staged tiles before
// the slot is released (the empty barrier counts two arrivals: the MMA commit and warp 3).
// The (tap, channel tile) units of one filter tile see the same dY tiles and share the work
// round-robin by frame chunk; each unit ends with 128 atomics.
#include "conv_umma.h"

namespace sl {

using namespace ptx;

namespace {

constexpr int BLOCK_M = 128;  // filters (co) per tile
constexpr int KT = 64;        // frames per pipeline stage
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;
constexpr int BOX_BYTES = KT * 128;          // one 64-frame x 64-channel box
constexpr int A_BYTES = 2 * BOX_BYTES;       // 128 filters
constexpr int LBO = BOX_BYTES;               // byte stride between 64-channel groups
constexpr int SBO = 1024;                    // byte stride between 8-frame groups
constexpr int STAGING_BYTES = BLOCK_M * 128; // 128 filters x 32 fp32 columns
constexpr int WARP_STAGING_BYTES = 32 * 128; // one epilogue warp's 32-row slab
constexpr int kEpiThreads = 128;

template <int BN>
struct WCfg {
  static constexpr int B_BYTES = (BN / 64) * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int kStages = BN >= 256 ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 2 * STAGING_BYTES + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ WgradParams p) {
  using C = WCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* staging = smem + C::kStages * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmDY);
    prefetch_tmap(&p.tmX);
    prefetch_tmap(&p.tmDW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2);  // tcgen05.commit of the MMA warp + the bias-gradient warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_s, C::TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();  // the previous kernel's outputs (dY, the zeroed / partially reduced dW) are visible from here

  // unit -> (tap fastest, then m_tile, n_tile, split): CTAs running together work on the
  // same K range, i.e. read the same dY / X slabs from L2, and reduce into different dW tiles.
  const int num_units = p.tap_units * p.m_tiles * p.n_tiles * p.ksplit;
  const int k_total = p.B * p.tchunks;  // (utterance, frame chunk) pairs
  const int k_per = (k_total + p.ksplit - 1) / p.ksplit;

  auto decode = [&](int unit, int& tap, int& mt, int& nt, int& k_begin, int& k_end) {
    int u = unit;
    tap = (u % p.tap_units) * p.tap_step;  // first tap of the unit (tap pairing: taps tap, tap + 1)
    u /= p.tap_units;
    mt = u % p.m_tiles;
    u /= p.m_tiles;
    nt = u % p.n_tiles;
    u /= p.n_tiles;
    const int sp = u;
    k_begin = sp * k_per;
    k_end = k_begin + k_per < k_total ? k_begin + k_per : k_total;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int issued = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        int tap, mt, nt, k_begin, k_end;
        decode(unit, tap, mt, nt, k_begin, k_end);
        const int jp = tap - p.pad_l;
        int q, par;
        if (p.stride == 1) {
          q = jp;
          par = 0;
        } else {
          q = jp >= 0 ? jp / p.stride : -((-jp + p.stride - 1) / p.stride);
          par = jp - q * p.stride;
        }
        // tap pairing: frame shift / parity of the second tap (a tap past the filter reads zeros
        // through an out-of-range channel-group coordinate)
        int q2 = 0, par2 = 0;
        const bool second_valid = p.tap_step == 2 && tap + 1 < p.taps;
        if (p.tap_step == 2) {
          const int jp2 = jp + 1;
          if (p.stride == 1) {
            q2 = jp2;
          } else {
            q2 = jp2 >= 0 ? jp2 / p.stride : -((-jp2 + p.stride - 1) / p.stride);
            par2 = jp2 - q2 * p.stride;
          }
        }
        const int co0 = mt * BLOCK_M;
        const int ci0 = nt * BN;
        for (int kk = k_begin; kk < k_end; ++kk) {
          const int b = kk / p.tchunks;
          const int tr = (kk - b * p.tchunks) * KT;
          for (int term = 0; term < p.terms; ++term) {
            const int dy_off = term == 2 ? p.dy_lo_off : 0;
            const int x_off = term == 1 ? p.x_lo_off : 0;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_s = smem + stage * C::STAGE_BYTES;
            uint8_t* b_s = a_s + A_BYTES;
            if (p.dbg_mode == 1 && issued >= C::kStages) {  // measurement aid: MMA on stale tiles
              mbar_arrive(&full_bar[stage]);
              if (++stage == C::kStages) {
                stage = 0;
                phase ^= 1;
              }
              continue;
            }
            ++issued;
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            if (p.grouped) {
              // one TMA per operand: the 64-channel groups are a tensor dimension of their own,
              // so a box {64 ch, 64 frames, n groups} lands as [group][frame][64 ch].  Groups past
              // the tensor read as zeros; for output_conv (64 filters) the second group of the
              // 128-row tile is zero or the lo plane, and those rows are clipped at the store.
              tma_load_4d(&p.tmDY, &full_bar[stage], a_s, 0, tr, (co0 + dy_off) >> 6, b);
              if (p.tap_step == 2) {
                // two boxes of 2 channel groups each: columns [0, 128) = tap, [128, 256) = tap + 1
                const int g = x_off >> 6;
                tma_load_5d(&p.tmX, &full_bar[stage], b_s, 0, par, tr + q, g, b);
                tma_load_5d(&p.tmX, &full_bar[stage], b_s + 2 * BOX_BYTES, 0, par2, tr + q2,
                            second_valid ? g : (1 << 20), b);
              } else {
                tma_load_5d(&p.tmX, &full_bar[stage], b_s, 0, par, tr + q, (ci0 + x_off) >> 6, b);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                // filter groups beyond cout_pad (output_conv: 64 < 128) read as zeros via an
                // out-of-range channel coordinate
                const int c = (co0 + 64 * i < p.cout_pad) ? (co0 + 64 * i + dy_off) : p.dy_c_total;
                tma_load_3d(&p.tmDY, &full_bar[stage], a_s + i * BOX_BYTES, c, tr, b);
              }
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_4d(&p.tmX, &full_bar[stage], b_s + i * BOX_BYTES, ci0 + 64 * i + x_off, par,
                            tr + q, b);
            }
            if (++stage == C::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_16(BLOCK_M, BN, 1, 1, p.fp16);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      int tap, mt, nt, k_begin, k_end;
      decode(unit, tap, mt, nt, k_begin, k_end);
      const int ksteps = (k_end - k_begin) * p.terms;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        // The issue loop is bound by this one thread's instruction stream, so the leader is
        // chosen with elect.sync — a single-taker branch ptxas can see, which keeps the operands in uniform
        // registers — and the descriptors of a stage differ only in their address field (one add each).
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t da0 = make_smem_desc_sw128(a_addr, LBO, SBO);
          const uint64_t db0 = make_smem_desc_sw128(a_addr + A_BYTES, LBO, SBO);
#pragma unroll
          for (int k = 0; k < KT / UMMA_K; ++k) {
            // 16 frames = two 8-row swizzle atoms = 2048 B further along K
            const uint64_t da = da0 + static_cast<uint64_t>(k * (2048 >> 4));
            const uint64_t db = db0 + static_cast<uint64_t>(k * (2048 >> 4));
            umma_bf16(tmem_d, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (ks == ksteps - 1) umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (ksteps == 0 && lane == 0) mbar_arrive(&tmem_full[as]);  // empty K range: nothing to add
    }
  } else if (warp == 3) {
    // ===================== bias gradient (column sums of the staged dY tiles) =====================
    int stage = 0;
    uint32_t phase = 0;
    const int grp = lane >> 4;          // which 64-filter half of the 128-filter tile
    const int sub = lane & 15;          // 4 consecutive filters: bytes [sub*8, sub*8+8) of the row
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      int tap, mt, nt, k_begin, k_end;
      decode(unit, tap, mt, nt, k_begin, k_end);
      // All taps x channel tiles of one (filter tile, K split) stream the same dY tiles, so the
      // column-sum work is dealt round-robin: this unit sums every `slices`-th frame chunk.
      const int slices = p.tap_units * p.n_tiles;
      const int my_slice = (tap / p.tap_step) * p.n_tiles + nt;
      const int ksteps = (k_end - k_begin) * p.terms;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      int term = 0;
      int kk_mod = k_begin % slices;  // (frame chunk index) mod slices
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        const bool sel = p.db != nullptr && kk_mod == my_slice;
        if (sel && term != 1) {  // terms 0 / 2 stage the hi / lo plane of dY; term 1 repeats hi
          const uint8_t* a_s = smem + stage * C::STAGE_BYTES + grp * BOX_BYTES;
#pragma unroll 8
          for (int r = 0; r < KT; ++r) {
            const int phys = (sub >> 1) ^ (r & 7);  // SWIZZLE_128B: 16-byte chunk index ^= row % 8
            const uint2 q = *reinterpret_cast<const uint2*>(a_s + r * 128 + phys * 16 + (sub & 1) * 8);
            const float2 f0 = unpack_16x2(q.x, p.fp16), f1 = unpack_16x2(q.y, p.fp16);
            acc[0] += f0.x;
            acc[1] += f0.y;
            acc[2] += f1.x;
            acc[3] += f1.y;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++term == p.terms) {
          term = 0;
          if (++kk_mod == slices) kk_mod = 0;
        }
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (p.db != nullptr) {
        const int co = mt * BLOCK_M + grp * 64 + sub * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (co + e < p.n_filters) atomicAdd(p.db + co + e, acc[e] * p.out_scale);
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 128;
    int it = 0;
    uint32_t store_count = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      int tap, mt, nt, k_begin, k_end;
      decode(unit, tap, mt, nt, k_begin, k_end);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const bool has_data = k_end > k_begin;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                             static_cast<uint32_t>(as * BN);
      if (has_data) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == BN / 32 - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
          }
          // per-warp 32-row slab: no cross-warp barrier (see conv_umma.cu)
          uint8_t* sbuf = staging + ew * (2 * WARP_STAGING_BYTES) + (store_count & 1) * WARP_STAGING_BYTES;
          if (lane == 0) tma_wait_group_read<1>();
          __syncwarp();
          uint8_t* rowp = sbuf + lane * 128;
          if (p.out_scale != 1.0f) {  // (uniform) undo the loss scale of the fp16 mode: exact, a power of two
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * p.out_scale);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int phys = j ^ (lane & 7);
            *reinterpret_cast<uint4*>(rowp + phys * 16) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // rows beyond cout_pad (output_conv: 64 of 128) are clipped by the tensor map
            // tap pairing: columns [128, 256) of the tile are the second tap's 128 input channels
            const int col = p.tap_step == 2 ? (c & 3) * 32 : nt * BN + c * 32;
            const int tap_c = p.tap_step == 2 ? tap + (c >> 2) : tap;
            if (tap_c < p.taps) {
              if (p.use_atomics)
                tma_reduce_add_3d(&p.tmDW, sbuf, col, mt * BLOCK_M + ew * 32, tap_c);
              else
                tma_store_3d(&p.tmDW, sbuf, col, mt * BLOCK_M + ew * 32, tap_c);
            }
            tma_commit_group();
          }
          ++store_count;
        }
      } else {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
      }
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN>
int launch(const WgradParams& p, int num_sms, cudaStream_t stream) {
  using C = WCfg<BN>;
  static unsigned long long configured = 0;  // bit d: attribute set for device d
  int dev = 0;
  SL_CUDA(cudaGetDevice(&dev));
  if (!((configured >> (dev & 63)) & 1ull)) {
    SL_CUDA(cudaFuncSetAttribute(wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 C::SMEM_BYTES));
    configured |= 1ull << (dev & 63);
  }
  const int num_units = p.tap_units * p.m_tiles * p.n_tiles * p.ksplit;
  const int grid = num_units < num_sms ? num_units : num_sms;
  SL_CUDA(launch_pdl(PDL_WGRAD, wgrad_kernel<BN>, dim3(grid), dim3(kThreads), C::SMEM_BYTES, stream, p));
  return 0;
}

}  // namespace

int wgrad_launch(const WgradParams& p, int block_n, int num_sms, cudaStream_t stream) {
  switch (block_n) {
    case 64:
      return launch<64>(p, num_sms, stream);
    case 128:
      return launch<128>(p, num_sms, stream);
    case 256:
      return launch<256>(p, num_sms, stream);
    default:
      set_error("wgrad: unsupported block_n");
      return 1;
  }
}

}  // namespace sl
