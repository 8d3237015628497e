// frontend.cu — audio -> z-normalised mel power-level spectrogram on the GPU.
//
// Replaces the librosa pipeline behind LabeledExample.z_normalized_transposed_spectrogram()
// (reference labeled_example.py:99-140): librosa.stft(n_fft=512, hop=128: periodic Hann window,
// center=True with reflect padding) -> |.|^2 -> 10 log10 floored at -150 dB -> mel projection of the
// dB values (Slaney filterbank, 128 bins) -> transpose -> global z-normalisation.  The output
// (B, T_max, 128) fp32, zero beyond each utterance's frame count, is exactly the zero-padded
// `input_batch` of net.py:583-585, so it feeds the tower without leaving HBM.
//
// One CTA (256 threads) per frame: windowed 512-point FFT in shared memory (9 radix-2 stages, one
// butterfly per thread per stage), power/dB, then 128 threads each take one mel bin.
#include "common.cuh"

namespace sl {

namespace {

constexpr int N_FFT = 512;
constexpr int HOP = 128;
constexpr int N_BINS = N_FFT / 2 + 1;  // 257
constexpr int N_MEL = 128;

__global__ void __launch_bounds__(256) spectrogram_kernel(const float* __restrict__ audio,
                                                           const int32_t* __restrict__ sample_counts,
                                                           const float* __restrict__ mel_t,  // (257, 128)
                                                           float* __restrict__ out,          // (B, T_max, 128)
                                                           int audio_stride, int T_max) {
  __shared__ float2 x[N_FFT];
  __shared__ float2 tw[N_FFT / 2];
  __shared__ float level[N_BINS + 3];
  const int b = blockIdx.y;
  const int t = blockIdx.x;
  const int n_samples = sample_counts[b];
  const int frames = 1 + n_samples / HOP;  // librosa center=True
  if (t >= frames) return;
  const int tid = threadIdx.x;
  const float* y = audio + static_cast<size_t>(b) * audio_stride;

  // twiddles e^{-2 pi i k / 512} and the windowed, reflect-padded frame in bit-reversed order
  {
    float s, c;
    sincospif(-2.0f * tid / N_FFT, &s, &c);
    tw[tid] = make_float2(c, s);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = tid + h * 256;
    int i = t * HOP + n - N_FFT / 2;  // index into the un-padded signal
    if (i < 0) i = -i;                // numpy 'reflect': no edge repeat
    if (i >= n_samples) i = 2 * (n_samples - 1) - i;
    i = i < 0 ? 0 : i;                // (signals shorter than the half window)
    const float w = 0.5f - 0.5f * cospif(2.0f * n / N_FFT);  // periodic Hann
    x[__brev(static_cast<unsigned>(n)) >> (32 - 9)] = make_float2(w * y[i], 0.f);
  }
  __syncthreads();
#pragma unroll
  for (int stage = 0; stage < 9; ++stage) {
    const int half = 1 << stage;
    const int j = tid & (half - 1);
    const int base = ((tid >> stage) << (stage + 1)) + j;
    const float2 w = tw[j << (8 - stage)];
    const float2 a = x[base], bb = x[base + half];
    const float2 wb = make_float2(w.x * bb.x - w.y * bb.y, w.x * bb.y + w.y * bb.x);
    x[base] = make_float2(a.x + wb.x, a.y + wb.y);
    x[base + half] = make_float2(a.x - wb.x, a.y - wb.y);
    __syncthreads();
  }
  // power level in dB, floored at -150 (0 -> -150): labeled_example.py:153-160
  for (int k = tid; k < N_BINS; k += 256) {
    const float p = x[k].x * x[k].x + x[k].y * x[k].y;
    float l = -150.f;
    if (p > 0.f) l = fmaxf(10.f * log10f(p), -150.f);
    level[k] = l;
  }
  __syncthreads();
  if (tid < N_MEL) {
    float acc = 0.f;
    for (int k = 0; k < N_BINS; ++k) acc = fmaf(mel_t[k * N_MEL + tid], level[k], acc);
    out[(static_cast<size_t>(b) * T_max + t) * N_MEL + tid] = acc;
  }
}

// per-utterance sum and sum of squares over the valid (frames x 128) block, in double
__global__ void moments_kernel(const float* __restrict__ x, const int32_t* __restrict__ frame_counts,
                               double* __restrict__ moments, int T_max, int F) {
  __shared__ double red[2][256];
  const int b = blockIdx.y;
  const size_t n = static_cast<size_t>(frame_counts[b]) * F;
  const float* xb = x + static_cast<size_t>(b) * T_max * F;
  double s = 0.0, q = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double v = xb[i];
    s += v;
    q += v * v;
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(&moments[2 * b], red[0][0]);
    atomicAdd(&moments[2 * b + 1], red[1][0]);
  }
}

// (x - mean) / std over the valid block (numpy population std), zeros beyond it
__global__ void normalize_kernel(float* __restrict__ x, const int32_t* __restrict__ frame_counts,
                                 const double* __restrict__ moments, int T_max, int F) {
  const int b = blockIdx.y;
  const size_t n = static_cast<size_t>(frame_counts[b]) * F;
  const size_t total = static_cast<size_t>(T_max) * F;
  const double mean = moments[2 * b] / static_cast<double>(n);
  const double var = moments[2 * b + 1] / static_cast<double>(n) - mean * mean;
  const float inv = static_cast<float>(1.0 / sqrt(var > 0 ? var : 1.0));
  const float m = static_cast<float>(mean);
  float* xb = x + static_cast<size_t>(b) * total;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    xb[i] = i < n ? (xb[i] - m) * inv : 0.f;
}

}  // namespace

int spectrogram_launch(const float* audio, const int32_t* sample_counts, const float* mel_t, float* out, int B,
                       int audio_stride, int T_max, cudaStream_t s) {
  dim3 grid(T_max, B);
  spectrogram_kernel<<<grid, 256, 0, s>>>(audio, sample_counts, mel_t, out, audio_stride, T_max);
  SL_CUDA(cudaGetLastError());
  return 0;
}

int z_normalize_launch(float* x, const int32_t* frame_counts, double* moments, int B, int T_max, int F,
                       cudaStream_t s) {
  SL_CUDA(cudaMemsetAsync(moments, 0, static_cast<size_t>(B) * 2 * sizeof(double), s));
  dim3 grid(32, B);
  moments_kernel<<<grid, 256, 0, s>>>(x, frame_counts, moments, T_max, F);
  SL_CUDA(cudaGetLastError());
  normalize_kernel<<<grid, 256, 0, s>>>(x, frame_counts, moments, T_max, F);
  SL_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sl
