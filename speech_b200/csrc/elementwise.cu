// Flat-buffer optimizer kernels: the tail of the reference's training step
// (train.py:32-35: clip_grad_norm(model.parameters(), 200) ; optimizer.step() with plain SGD,
// train.py:95-97) over ONE contiguous fp32 parameter buffer and ONE contiguous gradient buffer.
//   sb_sumsq        : sum of squares of the flat gradient (grid-stride, float4 loads, per-CTA
//                     partials combined in a fixed order: bit-reproducible across ranks) -> the
//                     global L2 norm used by the clip
//   sb_sgd_clip_step: p -= lr * min(1, max_norm / (norm + 1e-6)) * g   (momentum buffer optional),
//                     the clip coefficient is read from device memory: no host sync in the step;
//                     optionally also writes the bf16 copy of the updated parameters that the
//                     next step's tensor-core GEMMs / recurrence kernels take as operands, so no
//                     per-step cast kernels are needed
// Roofline: HBM (read g once for the norm; read p,g + write p for the update = 16 B/param total).
#include "common.cuh"

#include "../../include/speech_b200.h"

namespace sb {

// DETERMINISTIC: every CTA writes its partial sum to a scratch slot, and the last CTA to finish
// adds the slots in index order.  With atomics the summation order - hence the last bits of the
// norm, the clip coefficient and finally the parameters - differed between data-parallel ranks
// holding bit-identical gradients (bench.py dp_check: replicas drifted apart by 1e-5 in 70 steps).
static constexpr int SUMSQ_MAX_CTAS = 1024;
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n,
                                                    float* __restrict__ out,
                                                    float* __restrict__ partial,
                                                    unsigned int* __restrict__ ticket) {
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = __ldg(g4 + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = (n4 << 2) + blockIdx.x * 256LL + threadIdx.x; i < n;
       i += (long long)gridDim.x * 256)
    acc += g[i] * g[i];
  acc = warp_sum(acc);
  __shared__ float part[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    partial[blockIdx.x] = t;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // fixed-order reduction of the gridDim.x partials by the last CTA
  float s = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) s += __ldcg(partial + i);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    *out = t;
    *ticket = 0u;
  }
}

__global__ void __launch_bounds__(256)
sgd_clip_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom,
                     __nv_bfloat16* __restrict__ p16, long long n, const float* __restrict__ sumsq,
                     float lr, float momentum, float max_norm) {
  const float norm = sqrtf(*sumsq);
  const float coef = fminf(1.0f, max_norm / (norm + 1e-6f));   // torch.nn.utils.clip_grad_norm_
  const float scale = coef;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(mom);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 pv = p4[i];
    float4 gv = __ldg(g4 + i);
    gv.x *= scale; gv.y *= scale; gv.z *= scale; gv.w *= scale;
    if (mom) {
      float4 mv = m4[i];
      mv.x = momentum * mv.x + gv.x; mv.y = momentum * mv.y + gv.y;
      mv.z = momentum * mv.z + gv.z; mv.w = momentum * mv.w + gv.w;
      m4[i] = mv;
      gv = mv;
    }
    pv.x -= lr * gv.x; pv.y -= lr * gv.y; pv.z -= lr * gv.z; pv.w -= lr * gv.w;
    p4[i] = pv;
    if (p16)
      reinterpret_cast<uint2*>(p16)[i] = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
  }
  for (long long i = (n4 << 2) + blockIdx.x * 256LL + threadIdx.x; i < n;
       i += (long long)gridDim.x * 256) {
    float gv = g[i] * scale;
    if (mom) { mom[i] = momentum * mom[i] + gv; gv = mom[i]; }
    p[i] -= lr * gv;
    if (p16) p16[i] = __float2bfloat16_rn(p[i]);
  }
}

}  // namespace sb

using namespace sb;

extern "C" int sb_sumsq_workspace_size(size_t* bytes) {
  if (!bytes) return SB_ERR_INVALID;
  *bytes = (SUMSQ_MAX_CTAS + 4) * sizeof(float);
  return SB_OK;
}

// workspace: sb_sumsq_workspace_size bytes, ZERO-INITIALISED ONCE by the caller (the kernel leaves
// its ticket word at zero again); partial sums are combined in a fixed order: bit-reproducible.
extern "C" int sb_sumsq(const float* g, long long n, float* out, void* workspace, void* stream_) {
  if (!g || !out || !workspace || n <= 0) return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int grid = device_sm_count() * 4;
  if (grid > SUMSQ_MAX_CTAS) grid = SUMSQ_MAX_CTAS;
  float* partial = reinterpret_cast<float*>(workspace);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + SUMSQ_MAX_CTAS);
  sumsq_kernel<<<grid, 256, 0, stream>>>(g, n, out, partial, ticket);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_sgd_clip_step(float* params, const float* grads, float* momentum_buf,
                                void* params_bf16, long long n, const float* sumsq, float lr,
                                float momentum, float max_norm, void* stream_) {
  if (!params || !grads || !sumsq || n <= 0) return SB_ERR_INVALID;
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)momentum_buf) & 15) return SB_ERR_INVALID;
  if ((uintptr_t)params_bf16 & 7) return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = device_sm_count() * 4;
  sgd_clip_step_kernel<<<grid, 256, 0, stream>>>(params, grads, momentum_buf,
                                                 reinterpret_cast<__nv_bfloat16*>(params_bf16), n,
                                                 sumsq, lr, momentum, max_norm);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
