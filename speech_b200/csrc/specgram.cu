// sb_log_specgram: log power spectrogram + normalisation of int16 PCM, the featuriser in front of
// the encoder (reference: speech/loader.py:152-166 `log_specgram` = scipy.signal.spectrogram(
// window='hann', nperseg, noverlap, detrend=False) -> log(PSD + eps); loader.py:65-67
// `(x - mean) / std`).
//
// One CTA = 8 consecutive frames of one utterance x all nperseg/2+1 bins.  The frames' samples
// are converted, windowed (periodic Hann) and kept in shared memory as float64; thread k
// accumulates bin k of all 8 frames with a direct DFT in float64 (twiddles from a shared table
// indexed by (k*n) mod N, so there is no argument-reduction error): 2 table loads + 8 broadcast
// sample loads feed 16 DFMAs.  A direct DFT is O(N^2) but N = 320 (20 ms at 16 kHz): 6.6 G DFMA
// for a whole north-star batch (64 x 10 s), far below one training step, and it is exact to
// float64 rounding where the reference's complex64 FFT is not.
// The result is cast to float32 BEFORE +eps and log, as the reference does
// (`np.log(spec.T.astype(np.float32) + eps)`); frames past the utterance's end are written as 0,
// which is what the reference's zero_pad_concat pads the normalised features with.
//
// Roofline: nominally HBM (2 B in, 4*(N/2+1)/step B out per sample); practically FP64 / shared
// memory issue bound (direct DFT).
#include "common.cuh"

#include "../../include/speech_b200.h"

namespace sb {

static constexpr int SPEC_FPB = 8;        // frames per CTA
static constexpr int SPEC_MAX_N = 1024;   // nperseg limit (shared memory)

struct SpecParams {
  const short* pcm;
  const long long* offsets;   // [B] first sample of utterance b in pcm
  const int* n_samples;       // [B]
  const float* mean;          // [nbins] or null
  const float* stdev;         // [nbins] or null
  float* out;                 // [B][max_frames][nbins]
  double scale;               // 1 / (fs * sum(w^2))
  float eps;
  int nperseg, step, nbins, max_frames;
};

__global__ void __launch_bounds__(256)
log_specgram_kernel(const SpecParams p) {
  extern __shared__ double sm[];
  const int N = p.nperseg;
  double* tw_c = sm;                // [N] cos(2 pi t / N)
  double* tw_s = tw_c + N;          // [N] sin(2 pi t / N)
  double* xw = tw_s + N;            // [SPEC_FPB][N] windowed samples
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * SPEC_FPB;
  const int ns = p.n_samples[b];
  const int noverlap = N - p.step;
  const int n_frames = ns >= N ? (ns - noverlap) / p.step : 0;
  const short* x = p.pcm + p.offsets[b];

  if (f0 >= n_frames) {          // CTA entirely in the zero padding of a short utterance
    for (int i = threadIdx.x; i < SPEC_FPB * p.nbins; i += blockDim.x) {
      const int fr = f0 + i / p.nbins;
      if (fr < p.max_frames)
        p.out[((long long)b * p.max_frames + fr) * p.nbins + (i % p.nbins)] = 0.f;
    }
    return;
  }
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)n / (double)N, &s, &c);
    tw_c[n] = c;
    tw_s[n] = s;
  }
  for (int i = threadIdx.x; i < SPEC_FPB * N; i += blockDim.x) {
    const int f = i / N, n = i - f * N;
    const int fr = f0 + f;
    double v = 0.0;
    if (fr < n_frames) {
      double s, c;
      sincospi(2.0 * (double)n / (double)N, &s, &c);
      v = (double)x[(long long)fr * p.step + n] * (0.5 - 0.5 * c);   // periodic Hann
    }
    xw[i] = v;
  }
  __syncthreads();

  for (int k = threadIdx.x; k < p.nbins; k += blockDim.x) {
    double re[SPEC_FPB], im[SPEC_FPB];
#pragma unroll
    for (int f = 0; f < SPEC_FPB; ++f) { re[f] = 0.0; im[f] = 0.0; }
    int t = 0;
    for (int n = 0; n < N; ++n) {
      const double c = tw_c[t], s = tw_s[t];
#pragma unroll
      for (int f = 0; f < SPEC_FPB; ++f) {
        const double v = xw[f * N + n];
        re[f] = fma(v, c, re[f]);
        im[f] = fma(v, s, im[f]);
      }
      t += k;
      if (t >= N) t -= N;
    }
    // one-sided density: every bin except DC (and Nyquist when N is even) counts twice
    const bool twice = (k != 0) && !((N & 1) == 0 && k == N / 2);
    const double sc = twice ? 2.0 * p.scale : p.scale;
    const float mu = p.mean ? p.mean[k] : 0.f;
    const float sd = p.stdev ? p.stdev[k] : 1.f;
#pragma unroll
    for (int f = 0; f < SPEC_FPB; ++f) {
      const int fr = f0 + f;
      if (fr >= p.max_frames) break;
      float o = 0.f;
      if (fr < n_frames) {
        const float psd = (float)((re[f] * re[f] + im[f] * im[f]) * sc);
        o = (logf(psd + p.eps) - mu) / sd;
      }
      p.out[((long long)b * p.max_frames + fr) * p.nbins + k] = o;
    }
  }
}

}  // namespace sb

using namespace sb;

extern "C" int sb_log_specgram(const short* pcm, const long long* offsets, const int* n_samples,
                               int B, int nperseg, int step, double scale, float eps,
                               const float* mean, const float* stdev, float* out, int max_frames,
                               void* stream_) {
  if (!pcm || !offsets || !n_samples || !out || B <= 0 || max_frames <= 0) return SB_ERR_INVALID;
  if (nperseg < 2 || step < 1 || step > nperseg) return SB_ERR_INVALID;
  if (nperseg > SPEC_MAX_N || B > 65535) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SpecParams p;
  p.pcm = pcm; p.offsets = offsets; p.n_samples = n_samples; p.mean = mean; p.stdev = stdev;
  p.out = out; p.scale = scale; p.eps = eps; p.nperseg = nperseg; p.step = step;
  p.nbins = nperseg / 2 + 1; p.max_frames = max_frames;
  const size_t smem = (size_t)(2 + SPEC_FPB) * nperseg * sizeof(double);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    if (cudaFuncSetAttribute(log_specgram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return SB_ERR_CUDA;
    smem_set = smem;
  }
  dim3 grid((unsigned)((max_frames + SPEC_FPB - 1) / SPEC_FPB), (unsigned)B);
  int threads = (p.nbins + 31) / 32 * 32;     // one thread per bin (161 bins -> 6 warps)
  if (threads > 256) threads = 256;
  log_specgram_kernel<<<grid, threads, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
