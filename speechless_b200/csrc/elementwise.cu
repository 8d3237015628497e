// elementwise.cu — HBM-bound helper kernels: packing between the Keras-facing fp32 tensors
// and the packed 16-bit form (bf16, split bf16 or fp16), dropout, Keras-2 Adam.
#include "common.cuh"

namespace sl {

namespace {

// (B,T,C) fp32 -> (B,T_alloc,planes*c_pad) bf16; one thread per 8 channels: two 16-byte loads (when
// the rows are 16-byte aligned, i.e. C % 4 == 0), one 16-byte store per plane, 32-bit index arithmetic.
// (The first version took a pair of channels per thread with 64-bit divisions: 27 us for the 41 MB
// input batch of the bench shape, 1.5 TB/s.)
__global__ void pack_activation_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                       int B, int T, int C, int T_alloc, int c_pad, int planes, int fp16) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const unsigned groups = static_cast<unsigned>(c_pad) / 8u;
  const size_t total = static_cast<size_t>(B) * T_alloc * groups;
  const bool vec = (C & 3) == 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t row;
    unsigned g;
    if (total <= 0xffffffffull) {  // (uniform) 32-bit division
      const unsigned i32 = static_cast<unsigned>(i);
      const unsigned r32 = i32 / groups;
      row = r32;
      g = i32 - r32 * groups;
    } else {
      row = i / groups;
      g = static_cast<unsigned>(i - row * groups);
    }
    const unsigned b = static_cast<unsigned>(row / static_cast<unsigned>(T_alloc));
    const unsigned t = static_cast<unsigned>(row - static_cast<size_t>(b) * T_alloc);
    const int c = static_cast<int>(g) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (t < static_cast<unsigned>(T)) {
      const float* src = x + (static_cast<size_t>(b) * T + t) * C + c;
      if (vec && c + 8 <= C) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(src));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a0.x, v[1] = a0.y, v[2] = a0.z, v[3] = a0.w;
        v[4] = a1.x, v[5] = a1.y, v[6] = a1.z, v[7] = a1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c + e < C) v[e] = src[e];
      }
    }
    __nv_bfloat16* dst = y + row * (static_cast<size_t>(planes) * c_pad) + c;
    uint32_t hi[4], lo[4];
    if (fp16) {  // (uniform) one fp16 plane
#pragma unroll
      for (int e = 0; e < 4; ++e) hi[e] = pack_fp16x2(v[2 * e], v[2 * e + 1]);
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      continue;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      const __nv_bfloat162 l =
          __floats2bfloat162_rn(v[2 * e] - __bfloat162float(h.x), v[2 * e + 1] - __bfloat162float(h.y));
      hi[e] = *reinterpret_cast<const uint32_t*>(&h);
      lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(dst + c_pad) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void unpack_activation_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ x,
                                         int B, int T, int C, int T_alloc, int c_pad, int planes, int fp16) {
  const size_t total = static_cast<size_t>(B) * T * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t bt = i / C;
    const int t = static_cast<int>(bt % T);
    const int b = static_cast<int>(bt / T);
    const __nv_bfloat16* src =
        y + (static_cast<size_t>(b) * T_alloc + t) * (static_cast<size_t>(planes) * c_pad);
    float v = unpack_16(reinterpret_cast<const uint16_t*>(src)[c], fp16);
    if (planes == 2) v += __bfloat162float(src[c_pad + c]);
    x[i] = v;
  }
}

// Keras (k,Cin,Cout) <-> internal master (k,cout_pad,cin_pad), fp32.  Rarely called
// (weight load/save), so a plain gather is fine.
__global__ void keras_to_internal_kernel(const float* __restrict__ wk, float* __restrict__ wi, int k,
                                         int Cin, int Cout, int cin_pad, int cout_pad) {
  const size_t total = static_cast<size_t>(k) * cout_pad * cin_pad;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin_pad);
    const size_t r = i / cin_pad;
    const int co = static_cast<int>(r % cout_pad);
    const int j = static_cast<int>(r / cout_pad);
    wi[i] = (ci < Cin && co < Cout) ? wk[(static_cast<size_t>(j) * Cin + ci) * Cout + co] : 0.f;
  }
}
__global__ void internal_to_keras_kernel(const float* __restrict__ wi, float* __restrict__ wk, int k,
                                         int Cin, int Cout, int cin_pad, int cout_pad) {
  const size_t total = static_cast<size_t>(k) * Cin * Cout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cout);
    const size_t r = i / Cout;
    const int ci = static_cast<int>(r % Cin);
    const int j = static_cast<int>(r / Cin);
    wk[i] = wi[(static_cast<size_t>(j) * cout_pad + co) * cin_pad + ci];
  }
}

// internal fp32 (k,cout_pad,cin_pad) -> w_fwd bf16 (k,cout_pad,planes*cin_pad): same order
__global__ void pack_w_fwd_kernel(const float* __restrict__ wi, __nv_bfloat16* __restrict__ wf,
                                  size_t rows, int cin_pad, int planes, int fp16) {
  const size_t pairs = cin_pad / 2;
  const size_t total = rows * pairs;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / pairs;
    const int c = static_cast<int>(i - row * pairs) * 2;
    const float2 v = *reinterpret_cast<const float2*>(wi + row * cin_pad + c);
    __nv_bfloat16* dst = wf + row * (static_cast<size_t>(planes) * cin_pad);
    if (fp16) {  // (uniform)
      *reinterpret_cast<uint32_t*>(dst + c) = pack_fp16x2(v.x, v.y);
      continue;
    }
    const __nv_bfloat162 hi = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(dst + c) = hi;
    if (planes == 2)
      *reinterpret_cast<__nv_bfloat162*>(dst + cin_pad + c) =
          __floats2bfloat162_rn(v.x - __bfloat162float(hi.x), v.y - __bfloat162float(hi.y));
  }
}

// per-layer placement of the kernels inside the flat fp32 parameter buffer
struct AdamLayer {
  unsigned long long begin, end;  // float offsets of the layer's kernel [begin, end)
  __nv_bfloat16* w_fwd;           // (rows, planes * cin_pad) bf16 or null
  int cin_pad;
};
struct AdamLayers {
  AdamLayer layer[16];
  int count;
  int planes;
  int fp16;
};

// Keras-2 Adam over the whole flat buffer, fused with the refresh of the bf16
// tensor-core operands: the master layout (k, cout_pad, cin_pad) is the forward operand's
// layout, so each float4 of updated weights is re-emitted as bf16 (hi | lo planes) in place.
__global__ void adam_fused_kernel(float* __restrict__ p, const float* __restrict__ g,
                                  float* __restrict__ m, float* __restrict__ v, size_t n4, float lr_t,
                                  float b1, float b2, float eps, const __grid_constant__ AdamLayers L) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define SL_ADAM1(f)                            \
  mm.f = b1 * mm.f + (1.f - b1) * gg.f;        \
  vv.f = b2 * vv.f + (1.f - b2) * gg.f * gg.f; \
  pp.f = pp.f - lr_t * mm.f / (sqrtf(vv.f) + eps);
    SL_ADAM1(x) SL_ADAM1(y) SL_ADAM1(z) SL_ADAM1(w)
#undef SL_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    const unsigned long long idx = static_cast<unsigned long long>(i) * 4;
    for (int l = 0; l < L.count; ++l) {
      if (idx >= L.layer[l].begin && idx < L.layer[l].end) {
        if (L.layer[l].w_fwd != nullptr) {
          const unsigned long long local = idx - L.layer[l].begin;
          const int cin_pad = L.layer[l].cin_pad;
          const unsigned long long row = local / cin_pad;
          const int c = static_cast<int>(local - row * cin_pad);
          __nv_bfloat16* dst = L.layer[l].w_fwd + row * (static_cast<size_t>(L.planes) * cin_pad) + c;
          if (L.fp16) {  // (uniform)
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_fp16x2(pp.x, pp.y), pack_fp16x2(pp.z, pp.w));
            break;
          }
          const __nv_bfloat162 h0 = __floats2bfloat162_rn(pp.x, pp.y), h1 = __floats2bfloat162_rn(pp.z, pp.w);
          uint2 hv;
          hv.x = *reinterpret_cast<const uint32_t*>(&h0);
          hv.y = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(dst) = hv;
          if (L.planes == 2) {
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(pp.x - __bfloat162float(h0.x), pp.y - __bfloat162float(h0.y));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(pp.z - __bfloat162float(h1.x), pp.w - __bfloat162float(h1.y));
            uint2 lv;
            lv.x = *reinterpret_cast<const uint32_t*>(&l0);
            lv.y = *reinterpret_cast<const uint32_t*>(&l1);
            *reinterpret_cast<uint2*>(dst + cin_pad) = lv;
          }
        }
        break;
      }
    }
  }
}

// Keras-2 Adam: p -= lr_t * m / (sqrt(v) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr_t, float b1, float b2,
                            float eps) {
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define SL_ADAM1(f)                                   \
  mm.f = b1 * mm.f + (1.f - b1) * gg.f;               \
  vv.f = b2 * vv.f + (1.f - b2) * gg.f * gg.f;        \
  pp.f = pp.f - lr_t * mm.f / (sqrtf(vv.f) + eps);
    SL_ADAM1(x) SL_ADAM1(y) SL_ADAM1(z) SL_ADAM1(w)
#undef SL_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (size_t i = n4 * 4 + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// split-K dgrad finalize: fp32 sums (rows, c_pad) -> ReLU mask -> packed bf16 (hi | lo).
// One thread per 8 channels: two float4 loads, one mask byte, one (or two) 16-byte stores.
__global__ void dgrad_finalize_kernel(const float* __restrict__ acc, const uint8_t* __restrict__ mask,
                                      __nv_bfloat16* __restrict__ dx, size_t rows, int c_pad, int planes,
                                      int fp16, float out_scale) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int groups = c_pad / 8;
  const size_t total = rows * groups;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / groups;
    const int g = static_cast<int>(i - row * groups);
    const float4 a = *reinterpret_cast<const float4*>(acc + row * c_pad + g * 8);
    const float4 b = *reinterpret_cast<const float4*>(acc + row * c_pad + g * 8 + 4);
    float v[8] = {a.x * out_scale, a.y * out_scale, a.z * out_scale, a.w * out_scale,
                  b.x * out_scale, b.y * out_scale, b.z * out_scale, b.w * out_scale};
    if (mask != nullptr) {
      const unsigned m = mask[row * groups + g];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (!((m >> e) & 1u)) v[e] = 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = pack_16x2(v[2 * e], v[2 * e + 1], fp16);
      lo[e] = pack_bf16x2(v[2 * e] - bf16_round(v[2 * e]), v[2 * e + 1] - bf16_round(v[2 * e + 1]));
    }
    __nv_bfloat16* dst = dx + row * (static_cast<size_t>(planes) * c_pad) + g * 8;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(dst + c_pad) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Inverted dropout on a packed activation (keras.layers.Dropout in front of striding_conv /
// inner_conv_*, reference net.py:301-303): y = keep ? x / (1 - p) : 0, training phase only.
// One thread per 8 channels; the keep decision of element i is bit-reproducible from
// (seed, i): 16 random bits per element from a splitmix64 hash of the 4-element group index.
// mask_out bit = keep & (relu_mask_in bit, if given): exactly what the input-gradient epilogue
// of the consuming layer must multiply by (together with the 1/(1-p) scale).
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void dropout_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                               const uint8_t* __restrict__ relu_mask_in, uint8_t* __restrict__ mask_out, int B,
                               int T, int T_alloc, int c_pad, int planes, int fp16, unsigned threshold16,
                               float scale, unsigned long long seed) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int groups = c_pad / 8;
  const size_t total = static_cast<size_t>(B) * T_alloc * groups;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / groups;  // row of the (B, T_alloc) activation
    const int g = static_cast<int>(i - row * groups);
    const int b = static_cast<int>(row / T_alloc);
    const int t = static_cast<int>(row - static_cast<size_t>(b) * T_alloc);
    const size_t base = row * (static_cast<size_t>(planes) * c_pad) + g * 8;
    if (t >= T) {  // allocation padding rows stay zero
      *reinterpret_cast<uint4*>(y + base) = make_uint4(0u, 0u, 0u, 0u);
      if (planes == 2) *reinterpret_cast<uint4*>(y + base + c_pad) = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const unsigned long long r0 = splitmix64(seed ^ (2 * i)), r1 = splitmix64(seed ^ (2 * i + 1));
    unsigned keep = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (((r0 >> (16 * e)) & 0xffffu) >= threshold16) keep |= 1u << e;
      if (((r1 >> (16 * e)) & 0xffffu) >= threshold16) keep |= 1u << (4 + e);
    }
    const uint4 h = *reinterpret_cast<const uint4*>(x + base);
    uint4 l = make_uint4(0u, 0u, 0u, 0u);
    if (planes == 2) l = *reinterpret_cast<const uint4*>(x + base + c_pad);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
    uint32_t ho[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 hv = unpack_16x2(hw[e], fp16);
      float v0 = hv.x + __uint_as_float(lw[e] << 16);
      float v1 = hv.y + __uint_as_float(lw[e] & 0xffff0000u);
      v0 = ((keep >> (2 * e)) & 1u) ? v0 * scale : 0.f;
      v1 = ((keep >> (2 * e + 1)) & 1u) ? v1 * scale : 0.f;
      ho[e] = pack_16x2(v0, v1, fp16);
      lo[e] = pack_bf16x2(v0 - bf16_round(v0), v1 - bf16_round(v1));
    }
    *reinterpret_cast<uint4*>(y + base) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(y + base + c_pad) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    const size_t mrow = static_cast<size_t>(b) * T + t;  // masks are dense (B, T)
    unsigned m = keep;
    if (relu_mask_in != nullptr) m &= relu_mask_in[mrow * groups + g];
    mask_out[mrow * groups + g] = static_cast<uint8_t>(m);
  }
}

// Raw-wave input (reference net.py:310-312: Conv1D "wave_conv", k = 250, stride 160, SAME padding):
// the strided long filter becomes a plain GEMM once every output frame's receptive field is laid
// out as a row.  (B,T,C) fp32 -> (B,T_out,planes*c_pad) bf16 with row[t][j*C + c] = x[t*stride + j
// - pad_l][c] (zero outside [0,T) and for columns >= k*C), so that wave_conv runs through the same
// tcgen05 kernels as a 1-tap convolution over k*C input channels.  The training-phase Dropout in
// front of wave_conv (net.py:301-303) is applied here: the keep decision is a function of the
// *source* sample, so a dropped sample is dropped in every window that contains it.
__global__ void window_activation_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int B,
                                         int T, int C, int k, int stride, int T_out, int pad_l, int c_pad,
                                         int planes, int fp16, unsigned threshold16, float scale,
                                         unsigned long long seed) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const size_t total = static_cast<size_t>(B) * T_out * c_pad;
  const int kc = k * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % c_pad);
    const size_t row = i / c_pad;
    const int t = static_cast<int>(row % T_out);
    const int b = static_cast<int>(row / T_out);
    float v = 0.f;
    if (col < kc) {
      const int j = col / C, c = col - j * C;
      const int ts = t * stride + j - pad_l;
      if (ts >= 0 && ts < T) {
        const size_t src = (static_cast<size_t>(b) * T + ts) * C + c;
        v = x[src];
        if (threshold16 != 0) {
          const unsigned long long r = splitmix64(seed ^ (src >> 2));
          v = ((r >> (16 * (src & 3))) & 0xffffu) >= threshold16 ? v * scale : 0.f;
        }
      }
    }
    __nv_bfloat16* dst = y + row * (static_cast<size_t>(planes) * c_pad) + col;
    if (fp16) {  // (uniform)
      *reinterpret_cast<uint16_t*>(dst) = pack_16(v, 1);
      continue;
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    dst[0] = hi;
    if (planes == 2) dst[c_pad] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int pack_activation_launch(const float* x, void* y, int B, int T, int C, int T_alloc, int c_pad,
                           int prec, cudaStream_t s) {
  const int planes = prec_planes(prec);
  const size_t total = static_cast<size_t>(B) * T_alloc * (c_pad / 8);
  SL_CUDA(launch_pdl(PDL_ELEMENTWISE, pack_activation_kernel, dim3(grid_for(total, 256)), dim3(256), 0, s, x,
                     reinterpret_cast<__nv_bfloat16*>(y), B, T, C, T_alloc, c_pad, planes, prec_fp16(prec)));
  return 0;
}
int unpack_activation_launch(const void* y, float* x, int B, int T, int C, int T_alloc, int c_pad,
                             int prec, cudaStream_t s) {
  const int planes = prec_planes(prec);
  const size_t total = static_cast<size_t>(B) * T * C;
  unpack_activation_kernel<<<grid_for(total, 256), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(y), x, B, T, C, T_alloc, c_pad, planes, prec_fp16(prec));
  SL_CUDA(cudaGetLastError());
  return 0;
}
int keras_to_internal_launch(const float* wk, float* wi, int k, int Cin, int Cout, int cin_pad,
                             int cout_pad, cudaStream_t s) {
  const size_t total = static_cast<size_t>(k) * cout_pad * cin_pad;
  keras_to_internal_kernel<<<grid_for(total, 256), 256, 0, s>>>(wk, wi, k, Cin, Cout, cin_pad, cout_pad);
  SL_CUDA(cudaGetLastError());
  return 0;
}
int internal_to_keras_launch(const float* wi, float* wk, int k, int Cin, int Cout, int cin_pad,
                             int cout_pad, cudaStream_t s) {
  const size_t total = static_cast<size_t>(k) * Cin * Cout;
  internal_to_keras_kernel<<<grid_for(total, 256), 256, 0, s>>>(wi, wk, k, Cin, Cout, cin_pad, cout_pad);
  SL_CUDA(cudaGetLastError());
  return 0;
}
int pack_weights_internal_launch(const float* wi, void* wf, int k, int cin_pad, int cout_pad,
                                 int prec, cudaStream_t s) {
  const int planes = prec_planes(prec);
  const size_t rows = static_cast<size_t>(k) * cout_pad;
  pack_w_fwd_kernel<<<grid_for(rows * (cin_pad / 2), 256), 256, 0, s>>>(
      wi, reinterpret_cast<__nv_bfloat16*>(wf), rows, cin_pad, planes, prec_fp16(prec));
  SL_CUDA(cudaGetLastError());
  return 0;
}
int dgrad_finalize_launch(const float* acc, const void* mask, void* dx, size_t rows, int c_pad, int prec,
                          float out_scale, cudaStream_t s) {
  const int planes = prec_planes(prec);
  SL_CUDA(launch_pdl(PDL_ELEMENTWISE, dgrad_finalize_kernel, dim3(grid_for(rows * (c_pad / 8), 256)), dim3(256), 0, s, acc,
                     reinterpret_cast<const uint8_t*>(mask), reinterpret_cast<__nv_bfloat16*>(dx), rows, c_pad,
                     planes, prec_fp16(prec), out_scale));
  return 0;
}
int dropout_launch(const void* x, void* y, const void* relu_mask_in, void* mask_out, int B, int T, int T_alloc,
                   int c_pad, int prec, float p, unsigned long long seed, cudaStream_t s) {
  const int planes = prec_planes(prec);
  const unsigned threshold16 = static_cast<unsigned>(p * 65536.0f + 0.5f);
  const size_t rows = static_cast<size_t>(B) * T_alloc;
  dropout_kernel<<<grid_for(rows * (c_pad / 8), 256), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y),
      reinterpret_cast<const uint8_t*>(relu_mask_in), reinterpret_cast<uint8_t*>(mask_out), B, T, T_alloc, c_pad,
      planes, prec_fp16(prec), threshold16, 1.0f / (1.0f - static_cast<float>(threshold16) / 65536.0f), seed);
  SL_CUDA(cudaGetLastError());
  return 0;
}
int window_activation_launch(const float* x, void* y, int B, int T, int C, int k, int stride, int T_out, int pad_l,
                             int c_pad, int prec, float p, unsigned long long seed, cudaStream_t s) {
  const int planes = prec_planes(prec);
  const unsigned threshold16 = p > 0.f ? static_cast<unsigned>(p * 65536.0f + 0.5f) : 0u;
  const size_t total = static_cast<size_t>(B) * T_out * c_pad;
  window_activation_kernel<<<grid_for(total, 256), 256, 0, s>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), B, T, C, k, stride, T_out, pad_l, c_pad, planes, prec_fp16(prec),
      threshold16,
      1.0f / (1.0f - static_cast<float>(threshold16) / 65536.0f), seed);
  SL_CUDA(cudaGetLastError());
  return 0;
}
int adam_fused_launch(float* p, const float* g, float* m, float* v, size_t n, const size_t* begins,
                      const size_t* ends, void* const* w_fwd, const int* cin_pads, int n_layers,
                      int prec, float lr, float b1, float b2, float eps, int t, cudaStream_t s) {
  AdamLayers L;
  L.count = n_layers;
  L.planes = prec_planes(prec);
  L.fp16 = prec_fp16(prec);
  for (int i = 0; i < n_layers; ++i) {
    L.layer[i].begin = begins[i];
    L.layer[i].end = ends[i];
    L.layer[i].w_fwd = reinterpret_cast<__nv_bfloat16*>(w_fwd[i]);
    L.layer[i].cin_pad = cin_pads[i];
  }
  const double lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(b2), t)) /
                      (1.0 - pow(static_cast<double>(b1), t));
  SL_CUDA(launch_pdl(PDL_ELEMENTWISE, adam_fused_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, s, p, g, m, v, n / 4,
                     static_cast<float>(lr_t), b1, b2, eps, L));
  return 0;
}
int adam_launch(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2,
                float eps, int t, cudaStream_t s) {
  // lr_t in double on the host, as Keras computes it from python floats
  const double lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(b2), t)) /
                      (1.0 - pow(static_cast<double>(b1), t));
  adam_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, s>>>(p, g, m, v, n, static_cast<float>(lr_t), b1, b2,
                                                       eps);
  SL_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sl
