// elementwise.cu — HBM-bound helper kernels: packing between the Keras-facing fp32 tensors
// and the packed bf16 form, bias gradient, Keras-2 Adam.
#include "common.cuh"

namespace sl {

namespace {

// (B,T,C) fp32 -> (B,T_alloc,planes*c_pad) bf16; one thread per pair of channels
__global__ void pack_activation_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                       int B, int T, int C, int T_alloc, int c_pad, int planes) {
  const size_t pairs_per_row = c_pad / 2;
  const size_t total = static_cast<size_t>(B) * T_alloc * pairs_per_row;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / pairs_per_row;
    const int c = static_cast<int>(i - row * pairs_per_row) * 2;
    const int b = static_cast<int>(row / T_alloc);
    const int t = static_cast<int>(row - static_cast<size_t>(b) * T_alloc);
    float v0 = 0.f, v1 = 0.f;
    if (t < T) {
      const float* src = x + (static_cast<size_t>(b) * T + t) * C;
      if (c < C) v0 = src[c];
      if (c + 1 < C) v1 = src[c + 1];
    }
    __nv_bfloat16* dst = y + row * (static_cast<size_t>(planes) * c_pad);
    const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
    *reinterpret_cast<__nv_bfloat162*>(dst + c) = hi;
    if (planes == 2) {
      const __nv_bfloat162 lo =
          __floats2bfloat162_rn(v0 - __bfloat162float(hi.x), v1 - __bfloat162float(hi.y));
      *reinterpret_cast<__nv_bfloat162*>(dst + c_pad + c) = lo;
    }
  }
}

__global__ void unpack_activation_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ x,
                                         int B, int T, int C, int T_alloc, int c_pad, int planes) {
  const size_t total = static_cast<size_t>(B) * T * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t bt = i / C;
    const int t = static_cast<int>(bt % T);
    const int b = static_cast<int>(bt / T);
    const __nv_bfloat16* src =
        y + (static_cast<size_t>(b) * T_alloc + t) * (static_cast<size_t>(planes) * c_pad);
    float v = __bfloat162float(src[c]);
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
                                  size_t rows, int cin_pad, int planes) {
  const size_t pairs = cin_pad / 2;
  const size_t total = rows * pairs;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / pairs;
    const int c = static_cast<int>(i - row * pairs) * 2;
    const float2 v = *reinterpret_cast<const float2*>(wi + row * cin_pad + c);
    __nv_bfloat16* dst = wf + row * (static_cast<size_t>(planes) * cin_pad);
    const __nv_bfloat162 hi = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(dst + c) = hi;
    if (planes == 2)
      *reinterpret_cast<__nv_bfloat162*>(dst + cin_pad + c) =
          __floats2bfloat162_rn(v.x - __bfloat162float(hi.x), v.y - __bfloat162float(hi.y));
  }
}

// internal fp32 (k,cout_pad,cin_pad) -> w_dgrad bf16 (k,cin_pad,planes*cout_pad): per-tap
// transpose through a 32x33 smem tile (both sides coalesced)
__global__ void pack_w_dgrad_kernel(const float* __restrict__ wi, __nv_bfloat16* __restrict__ wd,
                                    int cin_pad, int cout_pad, int planes) {
  __shared__ float tile[32][33];
  const int j = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const float* src = wi + static_cast<size_t>(j) * cout_pad * cin_pad;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    tile[r][threadIdx.x] = src[static_cast<size_t>(co0 + r) * cin_pad + ci0 + threadIdx.x];
  __syncthreads();
  __nv_bfloat16* dst = wd + static_cast<size_t>(j) * cin_pad * (static_cast<size_t>(planes) * cout_pad);
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const float v = tile[threadIdx.x][r];  // (co = co0 + tx, ci = ci0 + r)
    __nv_bfloat16* row = dst + static_cast<size_t>(ci0 + r) * (static_cast<size_t>(planes) * cout_pad);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    row[co0 + threadIdx.x] = hi;
    if (planes == 2) row[cout_pad + co0 + threadIdx.x] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// db[c] (+)= sum over rows of dy (hi + lo planes).  blockDim = (32, 8): 32 channel pairs
// x 8 row lanes; grid = (c_pad/64, row_splits)
__global__ void bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db,
                                 size_t rows, int c_pad, int planes, int C) {
  __shared__ float red[8][64];
  const int c = blockIdx.x * 64 + threadIdx.x * 2;
  const size_t row_elems = static_cast<size_t>(planes) * c_pad;
  float s0 = 0.f, s1 = 0.f;
  for (size_t r = blockIdx.y * static_cast<size_t>(blockDim.y) + threadIdx.y; r < rows;
       r += static_cast<size_t>(gridDim.y) * blockDim.y) {
    const __nv_bfloat16* p = dy + r * row_elems + c;
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(p);
    s0 += __bfloat162float(h.x);
    s1 += __bfloat162float(h.y);
    if (planes == 2) {
      const __nv_bfloat162 l = *reinterpret_cast<const __nv_bfloat162*>(p + c_pad);
      s0 += __bfloat162float(l.x);
      s1 += __bfloat162float(l.y);
    }
  }
  red[threadIdx.y][threadIdx.x * 2] = s0;
  red[threadIdx.y][threadIdx.x * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float s = 0.f;
      for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x * 2 + e];
      if (c + e < C) atomicAdd(db + c + e, s);
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

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int pack_activation_launch(const float* x, void* y, int B, int T, int C, int T_alloc, int c_pad,
                           int planes, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * T_alloc * (c_pad / 2);
  pack_activation_kernel<<<grid_for(total, 256), 256, 0, s>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), B, T, C, T_alloc, c_pad, planes);
  SL_CUDA(cudaGetLastError());
  return 0;
}
int unpack_activation_launch(const void* y, float* x, int B, int T, int C, int T_alloc, int c_pad,
                             int planes, cudaStream_t s) {
  const size_t total = static_cast<size_t>(B) * T * C;
  unpack_activation_kernel<<<grid_for(total, 256), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(y), x, B, T, C, T_alloc, c_pad, planes);
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
int pack_weights_internal_launch(const float* wi, void* wf, void* wd, int k, int cin_pad,
                                 int cout_pad, int planes, cudaStream_t s) {
  if (wf != nullptr) {
    const size_t rows = static_cast<size_t>(k) * cout_pad;
    pack_w_fwd_kernel<<<grid_for(rows * (cin_pad / 2), 256), 256, 0, s>>>(
        wi, reinterpret_cast<__nv_bfloat16*>(wf), rows, cin_pad, planes);
    SL_CUDA(cudaGetLastError());
  }
  if (wd != nullptr) {
    dim3 grid(cin_pad / 32, cout_pad / 32, k), block(32, 8);
    pack_w_dgrad_kernel<<<grid, block, 0, s>>>(wi, reinterpret_cast<__nv_bfloat16*>(wd), cin_pad,
                                               cout_pad, planes);
    SL_CUDA(cudaGetLastError());
  }
  return 0;
}
int bias_grad_launch(const void* dy, float* db, size_t rows, int c_pad, int planes, int C,
                     cudaStream_t s) {
  size_t splits = (rows + 255) / 256;
  if (splits > 256) splits = 256;
  if (splits < 1) splits = 1;
  dim3 grid(c_pad / 64, static_cast<unsigned>(splits)), block(32, 8);
  bias_grad_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), db, rows,
                                          c_pad, planes, C);
  SL_CUDA(cudaGetLastError());
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
