// Conv2d(+ReLU) front-end of the encoder as im2col + tensor-core GEMM.
//
// Replaces the cuDNN convolutions behind nn.Conv2d in Model.__init__/encode
// (speech/models/model.py:19-29,60-71): valid (padding 0) 2-D convolution over (time, freq) with
// kernel (kh, kw) and stride s in both dims, followed by ReLU; the stack's output is flattened
// channel-major to (B, T', C*F') (model.py:66-71).
//
// Layouts (all row-major):
//   activations between layers: "pixel-major channels-last"  P[(b*To + t)*Fo + f][c]  f32 =
//       exactly the C matrix of the GEMM; ReLU is applied by whoever READS it (so the
//       pre-activation sign is available to the backward mask);
//   im2col matrix A[m][(i*kw + j)*Ci + ci] bf16 (K padded to a multiple of 8 with zeros); the
//       weights are permuted to the same K order on the host side (40 K elements);
//   conv = sb_gemm_bf16_tn(A, Wp) + bias  on tcgen05.
// Backward: dC (pre-activation grad, bf16) -> dW = dC^T A (GEMM on transposed copies, split-K),
//   dA = dC Wp (GEMM), col2im as a GATHER (each input pixel sums its <= ceil(kh/s)*ceil(kw/s)
//   taps; no atomics) fused with the ReLU mask of the layer below.
// All kernels here are HBM-bound elementwise/gather kernels (algorithmic bytes = one read of the
// source + one write of the destination); the FLOPs run in gemm.cu.
#include "common.cuh"

#include "../../include/speech_b200.h"

namespace sb {

typedef __nv_bfloat16 bf16;

// ---- im2col ------------------------------------------------------------------------------------
// src: P[(b*Ti + ti)*Fi + fi][Ci] f32 (relu on read if `relu`), dst: A[M][Kp] bf16
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int B, int Ti, int Fi, int Ci,
              int kh, int kw, int s, int To, int Fo, int Kp, int relu) {
  const long long M = (long long)B * To * Fo;
  const int K = kh * kw * Ci;
  const long long total = M * Kp;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * 256) {
    const long long m = idx / Kp;
    const int k = (int)(idx - m * Kp);
    float v = 0.f;
    if (k < K) {
      const int ci = k % Ci;
      const int ij = k / Ci;
      const int j = ij % kw, i = ij / kw;
      const int f = (int)(m % Fo);
      const long long bt = m / Fo;
      const int t = (int)(bt % To);
      const int b = (int)(bt / To);
      v = __ldg(src + (((long long)b * Ti + (s * t + i)) * Fi + (s * f + j)) * Ci + ci);
      if (relu) v = fmaxf(v, 0.f);
    }
    dst[idx] = __float2bfloat16_rn(v);
  }
}

// ---- final relayout: C[(b*To+t)*Fo+f][c] -> out[b][t][c*Fo + f] with ReLU ------------------------
__global__ void __launch_bounds__(256)
relu_to_bct_kernel(const float* __restrict__ C, float* __restrict__ out, int B, int To, int Fo,
                   int Co) {
  const long long total = (long long)B * To * Fo * Co;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * 256) {
    // idx enumerates the OUTPUT (f fastest) so that writes are coalesced
    const int f = (int)(idx % Fo);
    long long r = idx / Fo;
    const int c = (int)(r % Co);
    r /= Co;                       // r = b*To + t
    out[idx] = fmaxf(__ldg(C + (r * Fo + f) * Co + c), 0.f);
  }
}

// ---- top of backward: dY[b][t][c*Fo+f] * (C > 0) -> dC[(b*To+t)*Fo+f][c] bf16, db[c] += ----------
__global__ void __launch_bounds__(256)
dconv_top_kernel(const float* __restrict__ dY, const float* __restrict__ C, bf16* __restrict__ dC,
                 float* __restrict__ db, int B, int To, int Fo, int Co) {
  extern __shared__ float dbs[];
  for (int c = threadIdx.x; c < Co; c += 256) dbs[c] = 0.f;
  __syncthreads();
  const long long total = (long long)B * To * Fo * Co;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * 256) {
    const int c = (int)(idx % Co);     // idx enumerates dC (c fastest)
    const long long m = idx / Co;
    const int f = (int)(m % Fo);
    const long long r = m / Fo;
    float g = __ldg(dY + (r * Co + c) * Fo + f);
    if (__ldg(C + idx) <= 0.f) g = 0.f;
    dC[idx] = __float2bfloat16_rn(g);
    if (g != 0.f) atomicAdd(&dbs[c], g);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Co; c += 256)
    if (dbs[c] != 0.f) atomicAdd(db + c, dbs[c]);
}

// ---- col2im as a gather, fused with the ReLU mask of the layer below -----------------------------
// dA[m][(i*kw+j)*Ci + ci] f32 (ld = ldA) -> dCprev[(b*Ti+ti)*Fi+fi][ci] bf16 (masked by Pprev > 0)
__global__ void __launch_bounds__(256)
col2im_relu_kernel(const float* __restrict__ dA, long long ldA, const float* __restrict__ Pprev,
                   bf16* __restrict__ dCprev, float* __restrict__ db, int B, int Ti, int Fi,
                   int Ci, int kh, int kw, int s, int To, int Fo) {
  extern __shared__ float dbs[];
  for (int c = threadIdx.x; c < Ci; c += 256) dbs[c] = 0.f;
  __syncthreads();
  const long long total = (long long)B * Ti * Fi * Ci;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * 256) {
    const int ci = (int)(idx % Ci);
    const long long px = idx / Ci;
    const int fi = (int)(px % Fi);
    const long long bt = px / Fi;
    const int ti = (int)(bt % Ti);
    const int b = (int)(bt / Ti);
    float g = 0.f;
    if (__ldg(Pprev + idx) > 0.f) {
      for (int i = ti % s; i < kh; i += s) {
        const int t = (ti - i) / s;
        if (ti - i < 0 || t >= To) continue;
        for (int j = fi % s; j < kw; j += s) {
          const int f = (fi - j) / s;
          if (fi - j < 0 || f >= Fo) continue;
          const long long m = ((long long)b * To + t) * Fo + f;
          g += __ldg(dA + m * ldA + (i * kw + j) * Ci + ci);
        }
      }
    }
    dCprev[idx] = __float2bfloat16_rn(g);
    if (g != 0.f) atomicAdd(&dbs[ci], g);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Ci; c += 256)
    if (dbs[c] != 0.f) atomicAdd(db + c, dbs[c]);
}

// ---- bf16 matrix transpose [R][C] -> [C][Rp] (Rp >= R, leading dimension of the output) ---------
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long R, int C,
                      long long ld_src, long long ld_dst) {
  __shared__ bf16 tile[64][66];
  const long long r0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
  for (int rr = ty; rr < 64; rr += 4) {
    const long long r = r0 + rr;
    const int c = c0 + tx;
    tile[rr][tx] = (r < R && c < C) ? src[r * ld_src + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int cc = ty; cc < 64; cc += 4) {
    const int c = c0 + cc;
    const long long r = r0 + tx;
    if (c < C && r < R) dst[(long long)c * ld_dst + r] = tile[tx][cc];
  }
}

static int grid_for(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace sb

using namespace sb;

extern "C" int sb_conv_im2col(const float* src, void* dst_bf16, int B, int Ti, int Fi, int Ci,
                              int kh, int kw, int stride, int Kp, int relu, void* stream_) {
  if (!src || !dst_bf16 || B <= 0 || Ci <= 0 || kh <= 0 || kw <= 0 || stride <= 0)
    return SB_ERR_INVALID;
  const int To = (Ti - kh) / stride + 1, Fo = (Fi - kw) / stride + 1;
  if (To <= 0 || Fo <= 0 || Kp < kh * kw * Ci) return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)B * To * Fo * Kp;
  im2col_kernel<<<grid_for(total), 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst_bf16), B, Ti,
                                                     Fi, Ci, kh, kw, stride, To, Fo, Kp, relu);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_conv_relu_to_bct(const float* C, float* out, int B, int To, int Fo, int Co,
                                   void* stream_) {
  if (!C || !out || B <= 0 || To <= 0 || Fo <= 0 || Co <= 0) return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  relu_to_bct_kernel<<<grid_for((long long)B * To * Fo * Co), 256, 0, stream>>>(C, out, B, To, Fo,
                                                                               Co);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_conv_dtop(const float* dY, const float* C, void* dC_bf16, float* db, int B,
                            int To, int Fo, int Co, void* stream_) {
  if (!dY || !C || !dC_bf16 || !db || B <= 0 || To <= 0 || Fo <= 0 || Co <= 0)
    return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dconv_top_kernel<<<grid_for((long long)B * To * Fo * Co), 256, Co * sizeof(float), stream>>>(
      dY, C, reinterpret_cast<bf16*>(dC_bf16), db, B, To, Fo, Co);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_conv_col2im_relu(const float* dA, long long ldA, const float* Pprev,
                                   void* dCprev_bf16, float* db, int B, int Ti, int Fi, int Ci,
                                   int kh, int kw, int stride, void* stream_) {
  if (!dA || !Pprev || !dCprev_bf16 || !db || B <= 0) return SB_ERR_INVALID;
  const int To = (Ti - kh) / stride + 1, Fo = (Fi - kw) / stride + 1;
  if (To <= 0 || Fo <= 0) return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  col2im_relu_kernel<<<grid_for((long long)B * Ti * Fi * Ci), 256, Ci * sizeof(float), stream>>>(
      dA, ldA, Pprev, reinterpret_cast<bf16*>(dCprev_bf16), db, B, Ti, Fi, Ci, kh, kw, stride, To,
      Fo);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_transpose_bf16(const void* src, void* dst, long long R, int C, long long ld_src,
                                 long long ld_dst, void* stream_) {
  if (!src || !dst || R <= 0 || C <= 0) return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dim3 grid((unsigned)((R + 63) / 64), (unsigned)((C + 63) / 64));
  if (grid.y > 65535) return SB_ERR_UNSUPPORTED;
  transpose_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(src),
                                                  reinterpret_cast<bf16*>(dst), R, C, ld_src,
                                                  ld_dst);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
