// "Parity mode" forward recurrence: the GRU time loop in plain fp32 on CUDA cores.
//
// The production kernels (gru.cu) feed the tensor cores bf16 operands (fp32 accumulate), which
// moves the CTC loss by ~5e-6 relative at initialisation and more on weights with large recurrent
// gains.  This kernel is the reference-precision path ("keep an
// f32-accumulate parity mode and say which mode each number was taken in"): same interface,
// same time-major layout, fp32 weights and state, no rounding anywhere; the projections around it
// run as 3-pass split-bf16 GEMMs (ops.gemm_split).  It exists to MEASURE the bf16 path against,
// not to be fast: one launch per direction, 8 hidden units per CTA (24 rows of W_hh resident in
// shared memory as fp32), h_{t-1} read back from the fp32 state through L2 after a grid barrier.
// Replaces (in that mode) the cuDNN RNN behind nn.GRU, speech/models/model.py:35-39,73.
#include "common.cuh"
#include <cuda.h>

#include "../../include/speech_b200.h"

namespace sb {

static constexpr int F32_UPC = 8;      // hidden units per CTA
static constexpr int F32_KC = 128;     // K chunk of h staged in shared memory

struct GruF32Params {
  const float* gi;     // [T*Bp][ndir*3H]
  const float* whh;    // [ndir][3H][H] fp32
  const float* bhh;    // [ndir][3H]
  float* y;            // [T*Bp][ndir*H]
  unsigned int* ctr;   // grid-barrier counter of this direction (zero-initialised)
  int T, Bp, H, ndir, dir;
};

__global__ void __launch_bounds__(256, 1) gru_fwd_f32_kernel(const GruF32Params p) {
  extern __shared__ float f32_smem[];
  const int H = p.H, Bp = p.Bp, T = p.T, dir = p.dir;
  const int nC = gridDim.x;
  const int j0 = blockIdx.x * F32_UPC;
  const int tid = threadIdx.x;
  const int ld = F32_KC + 4;
  float* ws = f32_smem;                       // [24][H]   rows g*8 + u
  float* hs = ws + 24 * H;                    // [Bp][ld]
  const int D = p.ndir * H;
  for (int e = tid; e < 24 * (H / 4); e += 256) {
    const int r = e / (H / 4), c4 = e % (H / 4);
    const int g = r / F32_UPC, u = r % F32_UPC;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j0 + u < H)
      v = __ldg(reinterpret_cast<const float4*>(p.whh + ((long long)dir * 3 * H + g * H + j0 + u) * H) + c4);
    reinterpret_cast<float4*>(ws + r * H)[c4] = v;
  }
  __syncthreads();
  // thread -> (batch row b, unit pair q): rows b and b + 64, units j0 + 2q, j0 + 2q + 1
  const int b = tid & 63, q = tid >> 6;
  float hprev[2][2] = {{0.f, 0.f}, {0.f, 0.f}};        // [row half][unit]
  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : (T - 1 - step);
    const int tp = dir == 0 ? t - 1 : t + 1;
    float acc[2][6];
#pragma unroll
    for (int rh = 0; rh < 2; ++rh)
#pragma unroll
      for (int i = 0; i < 6; ++i) acc[rh][i] = 0.f;
    if (step > 0) {
      if (tid == 0) {
        unsigned int spins = 0;
        while (ld_acquire_gpu(p.ctr) < (unsigned int)nC * step)
          if (++spins > SB_SPIN_LIMIT) __trap();
      }
      __syncthreads();
      for (int k0 = 0; k0 < H; k0 += F32_KC) {
        const int kc = min(F32_KC, H - k0);
        __syncthreads();
        for (int e = tid; e < Bp * (kc / 4); e += 256) {
          const int r = e / (kc / 4), c4 = e % (kc / 4);
          reinterpret_cast<float4*>(hs + r * ld)[c4] = __ldcg(
              reinterpret_cast<const float4*>(p.y + ((long long)tp * Bp + r) * D + dir * H + k0) + c4);
        }
        __syncthreads();
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          const int row = b + 64 * rh;
          if (row >= Bp) continue;
          const float4* hr = reinterpret_cast<const float4*>(hs + row * ld);
          for (int c4 = 0; c4 < kc / 4; ++c4) {
            const float4 h = hr[c4];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const float4 w = reinterpret_cast<const float4*>(
                    ws + (g * F32_UPC + 2 * q + u) * H + k0)[c4];
                acc[rh][g * 2 + u] += w.x * h.x + w.y * h.y + w.z * h.z + w.w * h.w;
              }
          }
        }
      }
    }
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int row = b + 64 * rh;
      if (row >= Bp) continue;
      const long long m = (long long)t * Bp + row;
      const float* gi = p.gi + m * (p.ndir * 3 * H) + dir * 3 * H;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = j0 + 2 * q + u;
        if (j >= H) continue;
        const float* bh = p.bhh + dir * 3 * H;
        const float r = 1.f / (1.f + expf(-(gi[j] + acc[rh][0 + u] + bh[j])));
        const float z = 1.f / (1.f + expf(-(gi[H + j] + acc[rh][2 + u] + bh[H + j])));
        const float n = tanhf(gi[2 * H + j] + r * (acc[rh][4 + u] + bh[2 * H + j]));
        hprev[rh][u] = (1.f - z) * n + z * hprev[rh][u];
        p.y[m * D + dir * H + j] = hprev[rh][u];
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) red_release_gpu_add(p.ctr, 1u);
  }
}

}  // namespace sb

using namespace sb;

// gi / y as in sb_gru_fwd; whh_f32 [ndir][3H][H] fp32; barrier: ndir zero-initialised u32 words.
extern "C" int sb_gru_fwd_f32(const float* gi, const float* whh_f32, const float* bhh, float* y,
                              unsigned int* barrier, int T, int Bp, int H, int ndir, void* stream_) {
  if (!gi || !whh_f32 || !bhh || !y || !barrier) return SB_ERR_INVALID;
  if (T <= 0 || Bp <= 0 || Bp > 128 || H <= 0 || H % 4 != 0 || (ndir != 1 && ndir != 2))
    return SB_ERR_UNSUPPORTED;
  const int nC = (H + F32_UPC - 1) / F32_UPC;
  if (nC > device_sm_count()) return SB_ERR_UNSUPPORTED;      // all CTAs must be co-resident
  const size_t smem = ((size_t)24 * H + (size_t)Bp * (F32_KC + 4)) * sizeof(float);
  if (smem > 227 * 1024) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (cudaFuncSetAttribute(gru_fwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem) != cudaSuccess)
    return SB_ERR_CUDA;
  if (cudaMemsetAsync(barrier, 0, sizeof(unsigned int) * ndir, stream) != cudaSuccess)
    return SB_ERR_CUDA;
  for (int dir = 0; dir < ndir; ++dir) {
    GruF32Params p;
    p.gi = gi; p.whh = whh_f32; p.bhh = bhh; p.y = y; p.ctr = barrier + dir;
    p.T = T; p.Bp = Bp; p.H = H; p.ndir = ndir; p.dir = dir;
    void* args[] = {(void*)&p};
    if (cudaLaunchCooperativeKernel((const void*)gru_fwd_f32_kernel, dim3(nC), dim3(256), args, smem,
                                    stream) != cudaSuccess)
      return SB_ERR_CUDA;
  }
  return SB_OK;
}
