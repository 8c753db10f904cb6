// sb_attn_step: one decoder step of the additive location-aware attention, forward only.
//
// Replaces NNAttention.forward (speech/models/seq2seq.py:344-360) on the decode path
// (Seq2Seq.decode_step :114-137, used by infer :162-178 and beam_search :180-227):
//   score_t = w . relu(eh[t,:] + dhx + conv1d(ax_prev)[:, t]) + b      (conv: 1 -> H channels, 'same')
//   [score_t *= log T]  ;  ax = softmax_t(score)  ;  sx = sum_t ax_t eh[t,:]
// The reference materialises a (B,T,H) temporary and reads eh twice per step.  Here one CTA per
// utterance makes ONE pass over eh (coalesced rows, values kept in registers) with an online
// softmax: every warp keeps a running (max, sum, weighted row sum) over its frames, the 8 warps are
// merged at the end.  Algorithmic bytes = B*T*H*4 read (+ B*(T+H)*4 written) per step: HBM-bound.
#include "common.cuh"
#include <math.h>

#include "../../include/speech_b200.h"

namespace sb {

static constexpr int ATT_THREADS = 256;
static constexpr int ATT_MAXR = 32;   // H <= 32*32 = 1024

struct AttnParams {
  const float* eh;       // (B, T, H)
  const float* dhx;      // (B, H)
  const float* ax_prev;  // (B, T) or null
  const float* conv_w;   // (H, Kc)
  const float* conv_b;   // (H)
  const float* lin_w;    // (H)
  float lin_b;
  float* sx;             // (B, H)
  float* ax;             // (B, T)
  int B, T, H, Kc, log_t;
};

__global__ void __launch_bounds__(ATT_THREADS) attn_step_kernel(const AttnParams p) {
  extern __shared__ float att_smem[];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, H = p.H, Kc = p.Kc, pad = (Kc - 1) / 2;
  float* axp = att_smem;                       // [T + Kc - 1]  zero-padded previous alignment
  float* cw = axp + T + Kc - 1;                // [H * Kc]
  float* dc = cw + H * Kc;                     // [H] dhx + conv bias
  float* lw = dc + H;                          // [H]
  float* score = lw + H;                       // [T]
  float* wstat = score + T;                    // [8][2] per-warp (max, sum)
  float* wsx = wstat + 16;                     // [8][H] per-warp weighted sums
  const bool has_prev = p.ax_prev != nullptr;

  for (int k = tid; k < T + Kc - 1; k += ATT_THREADS) {
    const int t = k - pad;
    axp[k] = (has_prev && t >= 0 && t < T) ? p.ax_prev[(size_t)b * T + t] : 0.f;
  }
  if (has_prev)
    for (int k = tid; k < H * Kc; k += ATT_THREADS) cw[k] = p.conv_w[k];
  for (int h = tid; h < H; h += ATT_THREADS) {
    dc[h] = p.dhx[(size_t)b * H + h] + (has_prev ? p.conv_b[h] : 0.f);
    lw[h] = p.lin_w[h];
  }
  __syncthreads();

  const float tscale = p.log_t ? logf((float)T) : 1.0f;
  const float* eh = p.eh + (size_t)b * T * H;
  float m_run = -INFINITY, l_run = 0.f;
  float acc[ATT_MAXR];
#pragma unroll
  for (int r = 0; r < ATT_MAXR; ++r) acc[r] = 0.f;

  for (int t = warp; t < T; t += ATT_THREADS / 32) {
    float e[ATT_MAXR];
    float part = 0.f;
#pragma unroll
    for (int r = 0; r < ATT_MAXR; ++r) {
      const int h = lane + 32 * r;
      e[r] = 0.f;
      if (h < H) {
        e[r] = __ldg(eh + (size_t)t * H + h);
        float v = e[r] + dc[h];
        if (has_prev) {
          const float* c = cw + h * Kc;
          float s = 0.f;
          for (int k = 0; k < Kc; ++k) s += c[k] * axp[t + k];
          v += s;
        }
        part += lw[h] * fmaxf(v, 0.f);
      }
    }
    const float sc = (warp_sum(part) + p.lin_b) * tscale;
    if (lane == 0) score[t] = sc;
    const float m_new = fmaxf(m_run, sc);
    const float rescale = __expf(m_run - m_new);   // exp(-inf) = 0 on the first frame
    const float w = __expf(sc - m_new);
    l_run = l_run * rescale + w;
#pragma unroll
    for (int r = 0; r < ATT_MAXR; ++r) acc[r] = acc[r] * rescale + w * e[r];
    m_run = m_new;
  }
  if (lane == 0) { wstat[warp * 2] = m_run; wstat[warp * 2 + 1] = l_run; }
#pragma unroll
  for (int r = 0; r < ATT_MAXR; ++r) {
    const int h = lane + 32 * r;
    if (h < H) wsx[warp * H + h] = acc[r];
  }
  __syncthreads();
  float m = -INFINITY;
  for (int w = 0; w < ATT_THREADS / 32; ++w) m = fmaxf(m, wstat[w * 2]);
  float l = 0.f;
  for (int w = 0; w < ATT_THREADS / 32; ++w)
    l += (wstat[w * 2] == -INFINITY) ? 0.f : wstat[w * 2 + 1] * __expf(wstat[w * 2] - m);
  const float inv = 1.0f / l;
  for (int h = tid; h < H; h += ATT_THREADS) {
    float s = 0.f;
    for (int w = 0; w < ATT_THREADS / 32; ++w)
      if (wstat[w * 2] != -INFINITY) s += wsx[w * H + h] * __expf(wstat[w * 2] - m);
    p.sx[(size_t)b * H + h] = s * inv;
  }
  for (int t = tid; t < T; t += ATT_THREADS) p.ax[(size_t)b * T + t] = __expf(score[t] - m) * inv;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_attn_step(const float* eh, const float* dhx, const float* ax_prev,
                            const float* conv_w, const float* conv_b, const float* lin_w,
                            float lin_b, int log_t, int B, int T, int H, int Kc, float* sx,
                            float* ax, void* stream_) {
  if (!eh || !dhx || !conv_w || !conv_b || !lin_w || !sx || !ax) return SB_ERR_INVALID;
  if (B <= 0 || T <= 0 || H <= 0 || Kc <= 0 || (Kc & 1) == 0) return SB_ERR_INVALID;
  if (H > 32 * ATT_MAXR) return SB_ERR_UNSUPPORTED;
  const size_t smem = sizeof(float) * ((size_t)T + Kc - 1 + (size_t)H * Kc + 2 * H + T + 16 +
                                       (size_t)(ATT_THREADS / 32) * H);
  if (smem > 220 * 1024) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute(attn_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem) != cudaSuccess)
    return SB_ERR_CUDA;
  AttnParams p;
  p.eh = eh; p.dhx = dhx; p.ax_prev = ax_prev; p.conv_w = conv_w; p.conv_b = conv_b;
  p.lin_w = lin_w; p.lin_b = lin_b; p.sx = sx; p.ax = ax;
  p.B = B; p.T = T; p.H = H; p.Kc = Kc; p.log_t = log_t;
  attn_step_kernel<<<B, ATT_THREADS, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
