// frontend.cu — audio -> z-normalised mel power-level spectrogram on the GPU.
//
// Replaces the librosa pipeline behind LabeledExample.z_normalized_transposed_spectrogram()
// (reference labeled_example.py:99-140): librosa.stft(n_fft=512, hop=128: periodic Hann window,
// center=True with reflect padding) -> |.|^2 -> 10 log10 floored at -150 dB -> mel projection of the
// dB values (Slaney filterbank, 128 bins) -> transpose -> global z-normalisation.  The output
// (B, T_max, 128) fp32, zero beyond each utterance's frame count, is exactly the zero-padded
// `input_batch` of net.py:583-585, so it feeds the tower without leaving HBM.
//
// One CTA (256 threads) per FR consecutive frames of one utterance.  Per CTA, once: the audio span of
// its frames (coalesced, reflect-padded) -> smem; the Hann window and the FFT twiddles -> smem; the
// band [first, last] of non-zero weights of every mel filter (the Slaney filters are triangles a few
// bins wide: ~5 % of the dense 257 x 128 matrix) -> smem.  Per frame: windowed 512-point FFT in shared
// memory (9 radix-2 stages, one butterfly per thread per stage) of TWO real frames packed as one complex
// signal, power/dB of both, then 2 x 128 threads each take one mel bin of one frame over its band only.  (The first version did all of the set-up per frame and the
// dense projection: 1.07 ms for 64 x 1251 frames; this one: 0.26 ms)
#include "common.cuh"

namespace sl {

namespace {

constexpr int N_FFT = 512;
constexpr int HOP = 128;
constexpr int N_BINS = N_FFT / 2 + 1;  // 257
constexpr int N_MEL = 128;
constexpr int FR = 16;                 // frames per CTA
constexpr int SPAN = (FR - 1) * HOP + N_FFT;
constexpr int MAX_BAND = 40;           // widest supported mel filter, in FFT bins (dense fallback beyond)

__global__ void __launch_bounds__(256) spectrogram_kernel(const float* __restrict__ audio,
                                                           const int32_t* __restrict__ sample_counts,
                                                           const float* __restrict__ mel_t,  // (257, 128)
                                                           float* __restrict__ out,          // (B, T_max, 128)
                                                           int audio_stride, int T_max) {
  __shared__ float span[SPAN];
  __shared__ float window[N_FFT];
  // x is indexed through PX(i) = i + i / 32: one pad slot per 32 elements makes the bit-reversed scatter of
  // the windowed frame conflict free (without it the 32 lanes of a store hit one bank: ncu counted 58 % of
  // all shared-memory wavefronts of the first version as bank conflicts); tw holds the twiddles of every
  // stage contiguously (stage s, butterfly j at 2^s - 1 + j) instead of one strided table
  __shared__ float2 x[N_FFT + N_FFT / 32];
  __shared__ float2 tw[N_FFT];
  __shared__ float level[2][N_BINS + 3];
  __shared__ float band_w[N_MEL][MAX_BAND + 1];  // (+1: conflict-free column walks)
  __shared__ int band_first[N_MEL], band_len[N_MEL];
  const int b = blockIdx.y;
  const int t_begin = blockIdx.x * FR;
  const int n_samples = sample_counts[b];
  const int frames = 1 + n_samples / HOP;  // librosa center=True
  if (t_begin >= frames) return;
  const int t_end = min(t_begin + FR, frames);
  const int tid = threadIdx.x;
  const float* y = audio + static_cast<size_t>(b) * audio_stride;

  // ---- per-CTA set-up ----
  auto PX = [](int i) { return i + (i >> 5); };
  for (int idx = tid; idx < N_FFT - 1; idx += 256) {
    const int stage = 31 - __clz(idx + 1);  // entries 2^s - 1 .. 2^(s+1) - 2 belong to stage s
    const int j = idx + 1 - (1 << stage);
    float s, c;
    sincospif(-2.0f * (j << (8 - stage)) / N_FFT, &s, &c);  // e^{-2 pi i j 2^(8-s) / 512}
    tw[idx] = make_float2(c, s);
  }
  for (int n = tid; n < N_FFT; n += 256) window[n] = 0.5f - 0.5f * cospif(2.0f * n / N_FFT);  // periodic Hann
  for (int j = tid; j < SPAN; j += 256) {
    int i = t_begin * HOP + j - N_FFT / 2;  // index into the un-padded signal
    if (i < 0) i = -i;                      // numpy 'reflect': no edge repeat
    if (i >= n_samples) i = 2 * (n_samples - 1) - i;
    i = i < 0 ? 0 : i;                      // (signals shorter than the half window)
    span[j] = y[i];
  }
  // band of non-zero weights of every mel filter: all 256 threads walk the (257, 128) matrix once
  // (two threads per filter, every other row each; coalesced, L2 resident, independent loads — a
  // single thread per filter walking its column with a dependent branch per row cost more than the
  // FFTs of the CTA's 16 frames)
  if (tid < N_MEL) {
    band_first[tid] = N_BINS;
    band_len[tid] = -1;  // (holds the last non-zero row until the second barrier)
  }
  __syncthreads();
  {
    const int m = tid & (N_MEL - 1);
    int first = N_BINS, last = -1;
#pragma unroll 8
    for (int k = tid >> 7; k < N_BINS; k += 2) {
      const float w = __ldg(mel_t + k * N_MEL + m);
      if (w != 0.f) {
        first = min(first, k);
        last = max(last, k);
      }
    }
    if (last >= 0) {
      atomicMin(&band_first[m], first);
      atomicMax(&band_len[m], last);
    }
  }
  __syncthreads();
  if (tid < N_MEL) {
    const int first = band_first[tid], last = band_len[tid];
    const int len = last >= first ? last - first + 1 : 0;
    band_len[tid] = len;  // > MAX_BAND: dense fallback below
    if (len <= MAX_BAND)
      for (int j = 0; j < len; ++j) band_w[tid][j] = __ldg(mel_t + (first + j) * N_MEL + tid);
  }
  __syncthreads();

  // Two REAL frames per complex FFT: z = f_t + i f_{t+1}  =>  F_t[k] = (Z[k] + conj(Z[N-k])) / 2,
  // F_{t+1}[k] = (Z[k] - conj(Z[N-k])) / (2i).  An odd last frame is paired with zeros.
  for (int t = t_begin; t < t_end; t += 2) {
    const float* frame = span + (t - t_begin) * HOP;
    const bool pair = t + 1 < t_end;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = tid + h * 256;
      const float w = window[n];
      x[PX(static_cast<int>(__brev(static_cast<unsigned>(n)) >> (32 - 9)))] =
          make_float2(w * frame[n], pair ? w * frame[n + HOP] : 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int stage = 0; stage < 9; ++stage) {
      const int half = 1 << stage;
      const int j = tid & (half - 1);
      const int base = ((tid >> stage) << (stage + 1)) + j;
      const float2 w = tw[half - 1 + j];
      const int ia = PX(base), ib = PX(base + half);
      const float2 a = x[ia], bb = x[ib];
      const float2 wb = make_float2(w.x * bb.x - w.y * bb.y, w.x * bb.y + w.y * bb.x);
      x[ia] = make_float2(a.x + wb.x, a.y + wb.y);
      x[ib] = make_float2(a.x - wb.x, a.y - wb.y);
      __syncthreads();
    }
    // power level in dB, floored at -150 (0 -> -150): labeled_example.py:153-160
    for (int k = tid; k < N_BINS; k += 256) {
      const float2 z = x[PX(k)], zc = x[PX((N_FFT - k) & (N_FFT - 1))];
      const float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);  // F_t[k]
      const float br = 0.5f * (z.y + zc.y), bi = 0.5f * (zc.x - z.x);  // F_{t+1}[k]
      const float p0 = ar * ar + ai * ai, p1 = br * br + bi * bi;
      level[0][k] = p0 > 0.f ? fmaxf(10.f * log10f(p0), -150.f) : -150.f;
      level[1][k] = p1 > 0.f ? fmaxf(10.f * log10f(p1), -150.f) : -150.f;
    }
    __syncthreads();
    {
      const int which = tid >> 7, m = tid & (N_MEL - 1);  // threads 0..127: frame t, 128..255: frame t + 1
      if (which == 0 || pair) {
        const float* lv = level[which];
        const int first = band_first[m], len = band_len[m];
        float acc = 0.f;
        if (len <= MAX_BAND) {
          for (int j = 0; j < len; ++j) acc = fmaf(band_w[m][j], lv[first + j], acc);
        } else {
          for (int k = 0; k < N_BINS; ++k) acc = fmaf(mel_t[k * N_MEL + m], lv[k], acc);
        }
        out[(static_cast<size_t>(b) * T_max + t + which) * N_MEL + m] = acc;
      }
    }
    // (the next pair's writes to x / level happen behind its own barriers)
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
  dim3 grid((T_max + FR - 1) / FR, B);
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
