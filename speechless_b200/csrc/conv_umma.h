// conv_umma.h — parameter blocks of the tcgen05 implicit-GEMM kernels
#pragma once
#include "common.cuh"

namespace sl {

enum { EPI_PACKED = 0, EPI_SOFTMAX = 1, EPI_F32 = 2 };

// Forward / input-gradient implicit GEMM (conv_umma.cu)
struct ConvGemmParams {
  CUtensorMap tmA;  // activations, rank 4 {C_total, stride, T_alloc/stride, B}, box {64,1,128,1}
  CUtensorMap tmB;  // weights (k, cout, cin), rank 3 {cin_total, cout, taps}: box {64, BN, 1} (fwd: K-major B)
                    // or box {64, 64, 1} (dgrad: MN-major B, 64 cout rows of contraction)
  CUtensorMap tmAhalo;  // halo mode: as tmA with a box of 128 + taps - 1 frames (one load serves every tap)
  CUtensorMap tmBtail;  // as tmB with a box of BN / tail_split filters: the last partial wave's narrow tiles
  CUtensorMap tmY;  // packed bf16 output, rank 3 {C_total, T_out, B}, box {64,32,1}; EPI_F32: fp32 partial
                    // sums {C_pad, T_out, B}, box {32,32,1}, written with TMA reduce-add
  int B;
  int T_out;
  int m_tiles_per_utt;
  int n_tiles;
  // CTA pairs (cta_group::2): ctas = 2 runs the kernel as clusters of two CTAs that share one
  // 256-frame x 256-filter tile; each CTA stages its 128 frames of A and 128 filters of B (tmB then
  // has a box of BN / 2 filters) and the leader issues MMAs of M = 256 that read both halves
  int ctas;     // 1 or 2
  int m_units;  // ceil(B * m_tiles_per_utt / ctas): frame tiles per filter tile, in units of `ctas` tiles
  // tail splitting: the tiles of the last, partial wave of the persistent grid are cut into
  // tail_split (1, 2 or 4) narrower tiles so that the wave costs 1/tail_split of a tile time
  int full_tiles;  // tiles [0, full_tiles) are whole; the rest are split
  int tail_split;
  // tail K split (EPI_PACKED, single CTAs): a tcgen05.mma of M = 128 takes the same ~147 cycles for N = 64 as
  // for N = 256 (measured), so narrow tiles do not shorten the partial wave — splitting the contraction does.
  // Each tile of the partial wave becomes tail_ksplit work items over tail_per 64-channel chunks (all taps)
  // each.  The items that finish first add their fp32 partial sums into the tile's slot of a zero-filled
  // scratch buffer (TMA reduce-add); the item that draws the last ticket waits for them, adds the slot to
  // its own accumulator, writes zeros back (the buffer is all zeros again when the kernel ends) and runs
  // the normal epilogue.  tail_tickets / tail_done: one self-resetting counter pair per tile of the wave.
  int tail_ksplit;  // 1 = off
  int tail_per;
  CUtensorMap tmScratch;  // fp32 {BN, 128, tiles of the wave}, box {32, 32, 1}
  float* tail_scratch;
  unsigned* tail_tickets;
  unsigned* tail_done;
  // split K (EPI_F32 only): every tile is computed by ksplit work items, each over a contiguous
  // range of 64-channel chunks of the contraction (all taps), whose fp32 partial sums meet in HBM
  // through TMA reduce-add
  int ksplit;
  // halo mode (stride 1, taps > 1): the A operand of tap j is the same smem tile shifted by j
  // rows, so one halo tile of 128 + taps - 1 frames is loaded per 64-channel chunk and every tap's
  // MMA reads it through a descriptor whose start address is advanced by j * 128 bytes
  int halo;            // 0 = one A box per (tap, chunk)
  int halo_rows;       // 128 + taps - 1
  int halo_base_mode;  // bring-up: 0 = descriptor base_offset 0, 1 = (start address >> 7) & 7
  int taps;
  int chunks;  // 64-channel chunks of the contraction dimension
  int terms;   // 1 = one plane (bf16 or fp16), 3 = split bf16 (hi*hi + hi*lo + lo*hi)
  int fp16;    // operands and packed output are fp16 instead of bf16 (SL_PREC_FP16)
  int a_lo_off;
  int b_lo_off;
  int stride;
  int pad_l;
  // loop tap i reads activation frames shifted by i - pad_l and the weights of filter tap
  // w_tap0 + i * w_tap_step: forward (0, +1); input gradient (k-1, -1); input gradient of a
  // stride-2 layer, one output parity at a time (largest tap of that parity, -2)
  int w_tap0;
  int w_tap_step;
  const float* bias;
  int n_valid;  // number of real (unpadded) output channels: bias bound
  int relu;
  float out_scale;  // multiplies the (masked) output: 1/(1-p) of a dropout layer sitting below (dgrad), else 1
  int y_planes;  // 1 or 2 (hi | lo)
  int y_lo_off;
  // ReLU sign bitmask, 1 bit per (frame, channel), rows of mask_row_bytes = C_pad / 8 bytes:
  // written by the forward epilogue (y > 0), consumed by the dgrad epilogue of the layer above
  const uint8_t* mask_bits_in;
  uint8_t* mask_bits_out;
  int mask_row_bytes;
  // mask_bits_in row of output row t (tile space) = out_t_scale * t + out_t_off of mask_T rows per
  // utterance (stride-2 input gradient: tile rows are the frames of one parity)
  int mask_T;
  int out_t_scale;
  int out_t_off;
  // 1: walk the tiles from the last frame tile to the first.  A layer whose input is larger than L2 and was
  // written by the launch just before it (output_conv reading the 164 MB of big_conv_2) then starts on the
  // tiles that are still L2 resident instead of evicting them with the oldest ones
  int reverse_order;
  int dbg_mode;   // bring-up only (env SL_DBG_MODE): 1 = stop issuing TMA loads after the first ring fill
  // measurement aid (env SL_TIMELINE=1): 32 clock64() stamps per CTA — entry, setup done, first operands
  // landed, last MMA issued / accumulator complete / epilogue done of the CTA's first tiles, exit
  long long* timeline;
  int b_grouped;  // MN-major B: tmB is rank 4 {64, cout, cin_total/64, taps}, one box {64,64,BN/64,1} per stage
  float* probs;   // EPI_SOFTMAX outputs
  float* logits;
  float* logp;
  int V;
};

int conv_gemm_launch(const ConvGemmParams& p, int block_n, int epi, bool b_mn_major, int num_sms,
                     cudaStream_t stream);

// Weight-gradient GEMM (wgrad_umma.cu): both operands MN-major, contraction over time
struct WgradParams {
  CUtensorMap tmDY;  // rank 3 {Cout_total, T_out, B}, box {64, 64, 1}
  CUtensorMap tmX;   // rank 4 {Cin_total, stride, T_alloc/stride, B}, box {64, 1, 64, 1}
  CUtensorMap tmDW;  // fp32 rank 3 {cin_pad, cout_pad, taps}, box {32, 128, 1}: TMA (reduce-)store target
  float* dw;         // (taps, cout_pad, cin_pad) fp32
  float* db;         // (n_filters) fp32 bias gradient, accumulated with atomics; may be null
  int n_filters;     // real (unpadded) filter count
  int B;
  int T_out;
  int taps;
  int m_tiles;  // ceil(cout_pad / 128)
  int n_tiles;  // cin_pad / BN
  int ksplit;
  // halo mode (stride 1, taps > 1): the A operand of tap j is the same smem tile shifted by j
  // rows, so one halo tile of 128 + taps - 1 frames is loaded per 64-channel chunk and every tap's
  // MMA reads it through a descriptor whose start address is advanced by j * 128 bytes
  int halo;            // 0 = one A box per (tap, chunk)
  int halo_rows;       // 128 + taps - 1
  int halo_base_mode;  // bring-up: 0 = descriptor base_offset 0, 1 = (start address >> 7) & 7
  int tchunks;  // ceil(T_out / 64)
  int terms;
  int fp16;         // operands are fp16 instead of bf16 (SL_PREC_FP16)
  float out_scale;  // multiplies dW and db on their way out (1 / loss scale of the fp16 mode)
  int dy_lo_off;
  int x_lo_off;
  int stride;
  int pad_l;
  int cout_pad;
  int cin_pad;
  int dy_c_total;  // channel extent of the dY tensor map (an OOB coordinate for zero tiles)
  int use_atomics;  // 1: TMA reduce-add into dw (split K / accumulate), 0: plain TMA store
  int dbg_mode;     // bring-up only (env SL_DBG_MODE)
  int grouped;      // tmDY rank 4 {64, T_out, C/64, B} / tmX rank 5 {64, S, T/S, C/64, B}: one TMA per operand
  // tap pairing (cin_pad == 128, grouped maps only): a 256-column tile holds the 128 input channels of
  // TWO adjacent taps — the X box of tap j and the one of tap j + 1 — so the MMAs run at N = 256 (an
  // N = 128 MMA occupies the tensor pipe as long as an N = 256 one) and the dY tiles are staged once
  // per tap pair.  tap_step = 2, tap_units = ceil(taps / 2); otherwise tap_step = 1, tap_units = taps.
  int tap_step;
  int tap_units;
};

int wgrad_launch(const WgradParams& p, int block_n, int num_sms, cudaStream_t stream);

}  // namespace sl
