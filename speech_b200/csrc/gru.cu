// Persistent GRU recurrence kernels (forward and backward in time) for sm_100a.
//
// Replaces the cuDNN RNN reached through nn.GRU in the reference encoder
// (speech/models/model.py:35-39 construction, :73 call; gate order r,z,n; bidirectional halves
// are summed by the caller, model.py:75-77) and the pred-net GRU of the transducer
// (speech/models/transducer_model.py:23-26,68).
//
// The time-batched input projection gi = X W_ih^T + b_ih is done for all T at once by
// sb_gemm_bf16_tn; these kernels run the T-serial part.  One cooperative launch per layer covers
// BOTH directions:
//   * CTA c of direction d owns 16 hidden units j0..j0+15: its 48 rows of W_hh (r,z,n) stay
//     resident in shared memory (bf16, UMMA K-major SWIZZLE_128B chunks) for all T steps;
//   * per step the CTA all-gathers h_{t-1} (bf16, written by every CTA of the direction in the
//     previous step) from L2 into a 4-slot smem ring, and one thread issues tcgen05.mma
//     D[batch(128) x 48] += h_{t-1}[batch x 64] * Wslice[48 x 64]^T per K chunk, accumulating
//     in TMEM; loads and MMAs are pipelined through full/empty mbarriers;
//   * thread b (= TMEM lane = batch row) reads its 48 accumulators with tcgen05.ld, applies the
//     gate math in fp32 (h_{t-1} of its own units lives in registers across steps) and writes
//     h_t as fp32 (Y), bf16 (next GEMM operand) and bf16-transposed (wgrad operand);
//   * a per-direction grid barrier (red.release / ld.acquire on a global counter) separates steps.
// The backward kernel has the same structure with W_hh^T resident (16 rows x 3H) and the
// all-gather over the pre-activation gradients dgh_t (batch x 3H).
//
// Roofline: the MMAs are tensor work but each step is bound by the all-gather + barrier latency;
// DESIGN.md reports us/step next to the tensor-pipe share.
#include "common.cuh"
#include <cuda.h>

#include "../../include/speech_b200.h"

namespace sb {

static constexpr int GRU_HC = 16;           // hidden units per CTA
static constexpr int GRU_MAX_RING = 16;     // smem ring slots for the gathered operand
static constexpr int GRU_EPI = 128;         // warps 0..3: epilogue (thread = TMEM lane = batch row)
static constexpr int GRU_THREADS = GRU_EPI + 64;  // + warp 4: MMA issuer/TMEM owner, warp 5: TMA

typedef __nv_bfloat16 bf16;

struct GruFwdParams {
  const float* gi;     // [T*Bp][ndir*3H]  input projections (b_ih already added)
  const bf16* whh;     // [ndir][3H][H]    recurrent weights, bf16
  const float* bhh;    // [ndir][3H]
  float* y;            // [T*Bp][ndir*H]   h_t fp32
  bf16* xn;            // [T*Bp][ndir*H]   h_t bf16 (operand of the next projection)
  bf16* xnT;           // [ndir*H][(T+2)*Bp] h_t bf16 transposed, column (t+1)*Bp+b ; may be null
  float* gates;        // [T*Bp][ndir][4][H] saved r,z,n,hn for backward ; may be null
  unsigned int* barrier;  // [ndir] zero-initialised counters
  unsigned long long* dbg;  // optional timeline (CTA 0): [step][16] globaltimer stamps, or null
  int T, Bp, H, ndir, ring;
};

struct GruBwdParams {
  const float* dy;     // [T*Bp][ndir*H]   gradient w.r.t. this layer's output
  const float* y;      // [T*Bp][ndir*H]   forward h_t (fp32)
  const float* gates;  // [T*Bp][ndir][4][H]
  const bf16* whhT;    // [ndir][H][3H]    W_hh^T, bf16
  bf16* dgi;           // [T*Bp][ndir*3H]  d(pre-activation) of the input projection, bf16
  bf16* dgiT;          // [ndir*3H][T*Bp]  same, transposed (wgrad operand)
  bf16* dghnT;         // [ndir][H][T*Bp]  dn_pre * r, transposed (wgrad of W_hn)
  bf16* xchg;          // [ndir][2][Bp][3H] per-step exchange of dgh_t
  float* dbih;         // [ndir*3H] += sum_{t,b} dgi
  float* dbhh;         // [ndir*3H] += sum_{t,b} dgh
  unsigned int* barrier;  // [ndir]
  int T, Bp, H, ndir, ring;
};

SB_DEVINL unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GRU_STAMP(ev)                                                          \
  do {                                                                         \
    if (p.dbg && blockIdx.x == 0 && step < 64) p.dbg[step * 16 + (ev)] = gtime(); \
  } while (0)

// ---- per-direction grid barrier ------------------------------------------------------------
SB_DEVINL void grid_arrive(unsigned int* ctr) { red_release_gpu_add(ctr, 1u); }
SB_DEVINL void grid_wait(const unsigned int* ctr, unsigned int target) {
  unsigned int spins = 0;
  while (ld_acquire_gpu(ctr) < target) {
    if (++spins > SB_SPIN_LIMIT) __trap();
  }
}
SB_DEVINL void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(GRU_EPI) : "memory"); }

struct GruSmem {
  uint8_t* wtile;   // resident weight chunks
  uint8_t* ring;    // ring slots, stride = Bp*128 bytes (+ slack so a 128-row read stays inside)
  uint64_t* full;   // [GRU_MAX_RING]
  uint64_t* empty;  // [GRU_MAX_RING]
  uint64_t* accfull;
  uint32_t* tmem_slot;
  float* scratch;   // [64]
};

SB_DEVINL GruSmem carve(uint8_t* raw, int wbytes, int ring_bytes) {
  GruSmem s;
  s.wtile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) &
                                       ~static_cast<uintptr_t>(1023));
  s.ring = s.wtile + wbytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.ring + ring_bytes);
  s.full = bars;
  s.empty = bars + GRU_MAX_RING;
  s.accfull = bars + 2 * GRU_MAX_RING;
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GRU_MAX_RING + 1);
  s.scratch = reinterpret_cast<float*>(bars + 2 * GRU_MAX_RING + 2);
  return s;
}

// ring geometry shared by host and device
SB_DEVINL int ring_stride(int Bp) { return Bp * 128; }

// TMA producer (one thread): bring `nchunks` [Bp x 64] bf16 boxes of the operand whose rows start
// at `row0` into the ring.  `fill` is the running chunk counter shared (by construction) with
// the MMA thread.
SB_DEVINL void tma_gather(const GruSmem& s, const CUtensorMap* tm, int row0, int Bp, int nchunks,
                          int ring, unsigned int& fill) {
  const int stride = ring_stride(Bp);
  for (int c = 0; c < nchunks; ++c) {
    const unsigned int slot = fill % ring;
    const unsigned int par = (fill / ring) & 1u;
    mbar_wait(&s.empty[slot], par ^ 1u);
    mbar_expect_tx(&s.full[slot], (uint32_t)stride);
    tma_load_2d(s.ring + slot * stride, tm, &s.full[slot], c * 64, row0);
    ++fill;
  }
}

// MMA thread: consume `nchunks` ring slots against the resident weight chunks.
template <int N>
SB_DEVINL void mma_consume(const GruSmem& s, uint32_t tmem_d, int nchunks, int wchunk_bytes,
                           int Bp, int ring, unsigned int& fill) {
  constexpr uint32_t idesc = umma_idesc_bf16_f32(128, N);
  const int stride = ring_stride(Bp);
  for (int c = 0; c < nchunks; ++c) {
    const unsigned int slot = fill % ring;
    const unsigned int par = (fill / ring) & 1u;
    mbar_wait(&s.full[slot], par);
    tc_fence_after_sync();
    const uint64_t da = umma_desc_sw128_kmajor(smem_u32(s.ring + slot * stride));
    const uint64_t db = umma_desc_sw128_kmajor(smem_u32(s.wtile + c * wchunk_bytes));
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                   (c > 0 || k > 0) ? 1u : 0u);
    umma_commit(&s.empty[slot]);
    ++fill;
  }
  umma_commit(s.accfull);
}

// =============================================================================================
// forward
// =============================================================================================
__global__ void __launch_bounds__(GRU_THREADS, 1)
gru_fwd_kernel(const __grid_constant__ CUtensorMap tm_d0, const __grid_constant__ CUtensorMap tm_d1,
               const GruFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int H = p.H, Bp = p.Bp, T = p.T;
  const int nC = H / GRU_HC;
  const int dir = blockIdx.x / nC;
  const int j0 = (blockIdx.x % nC) * GRU_HC;
  const int nchunks = (H + 63) / 64;
  constexpr int WCHUNK = 48 * 128;  // 48 rows x 64 bf16
  const int ring_bytes = p.ring * ring_stride(Bp) + (128 - Bp) * 128;
  const GruSmem s = carve(smem_raw, nchunks * WCHUNK, ring_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = p.ndir * H;
  const long long ldT = (long long)(T + 2) * Bp;
  const CUtensorMap* tm = dir == 0 ? &tm_d0 : &tm_d1;

  // ---- one-time setup: zero the ring, stage this CTA's 48 weight rows, barriers, TMEM ----
  for (int k = tid; k < ring_bytes / 16; k += GRU_THREADS)
    reinterpret_cast<uint4*>(s.ring)[k] = make_uint4(0, 0, 0, 0);
  {
    const int pieces_per_row = nchunks * 8;
    for (int k = tid; k < 48 * pieces_per_row; k += GRU_THREADS) {
      const int r = k / pieces_per_row, pc = k % pieces_per_row;
      const int g = r / GRU_HC, jj = r % GRU_HC;
      const int col = pc * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (col < H)
        v = *reinterpret_cast<const uint4*>(p.whh + ((long long)dir * 3 * H + g * H + j0 + jj) * H +
                                            col);
      *reinterpret_cast<uint4*>(s.wtile + (pc >> 3) * WCHUNK + sw128_offset(r, pc & 7)) = v;
    }
    if (tid < 48) {
      const int g = tid / GRU_HC, jj = tid % GRU_HC;
      s.scratch[tid] = p.bhh[dir * 3 * H + g * H + j0 + jj];
    }
  }
  if (tid == 0) {
    for (int i = 0; i < GRU_MAX_RING; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(s.accfull, 1);
    mbar_fence_init();
    tma_prefetch_desc(tm);
  }
  if (warp == 4) tmem_alloc(s.tmem_slot, 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s.tmem_slot;
  unsigned int* ctr = p.barrier + dir;

  if (warp == 5) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      unsigned int fill = 0;
      for (int step = 1; step < T; ++step) {
        const int t = dir == 0 ? step : (T - 1 - step);
        const int tp = dir == 0 ? t - 1 : t + 1;
        grid_wait(ctr, (unsigned int)nC * step);   // every CTA of this direction published h_{tp}
        GRU_STAMP(0);
        fence_proxy_async_all();                    // generic-proxy writes -> async-proxy (TMA) reads
        tma_gather(s, tm, tp * Bp, Bp, nchunks, p.ring, fill);
        GRU_STAMP(1);
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      unsigned int fill = 0;
      for (int step = 1; step < T; ++step) {
        mma_consume<48>(s, tmem_base, nchunks, WCHUNK, Bp, p.ring, fill);
        GRU_STAMP(2);
      }
    }
  } else {
    // ===================== epilogue: thread = batch row =====================
    float hprev[GRU_HC];
#pragma unroll
    for (int jj = 0; jj < GRU_HC; ++jj) hprev[jj] = 0.f;
    float bias[48];
#pragma unroll
    for (int r = 0; r < 48; ++r) bias[r] = s.scratch[r];
    const bool active = tid < Bp;
    for (int step = 0; step < T; ++step) {
      const int t = dir == 0 ? step : (T - 1 - step);
      // this step's input projections do not depend on the recurrence: fetch them first
      float gi[48];
      if (active) {
        const float* g = p.gi + ((long long)t * Bp + tid) * (p.ndir * 3 * H) + dir * 3 * H + j0;
#pragma unroll
        for (int gg = 0; gg < 3; ++gg)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(g + gg * H) + q);
            gi[gg * 16 + q * 4 + 0] = v.x; gi[gg * 16 + q * 4 + 1] = v.y;
            gi[gg * 16 + q * 4 + 2] = v.z; gi[gg * 16 + q * 4 + 3] = v.w;
          }
      }
      float acc[48];
      if (step > 0) {
        mbar_wait(s.accfull, (step - 1) & 1);
        if (tid == 0) GRU_STAMP(3);
        tc_fence_after_sync();
        uint32_t v[16];
#pragma unroll
        for (int gg = 0; gg < 3; ++gg) {
          tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + gg * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) acc[gg * 16 + jj] = __uint_as_float(v[jj]);
        }
        tc_fence_before_sync();
        if (tid == 0) GRU_STAMP(4);
      } else {
#pragma unroll
        for (int r = 0; r < 48; ++r) acc[r] = 0.f;
      }
      const long long m = (long long)t * Bp + tid;
      float hn[GRU_HC], rr[GRU_HC], zz[GRU_HC], nn[GRU_HC];
      if (active) {
#pragma unroll
        for (int jj = 0; jj < GRU_HC; ++jj) {
          rr[jj] = sigmoidf_fast(gi[jj] + acc[jj] + bias[jj]);
          zz[jj] = sigmoidf_fast(gi[16 + jj] + acc[16 + jj] + bias[16 + jj]);
          hn[jj] = acc[32 + jj] + bias[32 + jj];
          nn[jj] = tanhf_fast(gi[32 + jj] + rr[jj] * hn[jj]);
          hprev[jj] = (1.f - zz[jj]) * nn[jj] + zz[jj] * hprev[jj];
        }
        // critical path: only the bf16 h_t that the other CTAs gather next step
        uint4 pk[2];
        uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
        for (int q = 0; q < 8; ++q) pw[q] = pack_bf16x2(hprev[2 * q], hprev[2 * q + 1]);
        uint4* xo = reinterpret_cast<uint4*>(p.xn + m * D + dir * H + j0);
        xo[0] = pk[0];
        xo[1] = pk[1];
        if (tid == 0) GRU_STAMP(5);
        fence_proxy_async_all();   // these generic-proxy writes are read by other CTAs' TMA
        if (tid == 0) GRU_STAMP(6);
      }
      epi_barrier();
      if (tid == 0) {
        GRU_STAMP(7);
        __threadfence();
        GRU_STAMP(8);
        grid_arrive(ctr);
        GRU_STAMP(9);
      }
      // off the critical path: fp32 state, transposed copy, saved gates
      if (active) {
        float* yo = p.y + m * D + dir * H + j0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(yo)[q] =
              make_float4(hprev[q * 4], hprev[q * 4 + 1], hprev[q * 4 + 2], hprev[q * 4 + 3]);
        if (p.xnT) {
          bf16* xt = p.xnT + (long long)(dir * H + j0) * ldT + (long long)(t + 1) * Bp + tid;
#pragma unroll
          for (int jj = 0; jj < GRU_HC; ++jj) xt[jj * ldT] = __float2bfloat16_rn(hprev[jj]);
        }
        if (p.gates) {
          float* go = p.gates + ((m * p.ndir + dir) * 4) * H + j0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            reinterpret_cast<float4*>(go)[q] =
                make_float4(rr[q * 4], rr[q * 4 + 1], rr[q * 4 + 2], rr[q * 4 + 3]);
            reinterpret_cast<float4*>(go + H)[q] =
                make_float4(zz[q * 4], zz[q * 4 + 1], zz[q * 4 + 2], zz[q * 4 + 3]);
            reinterpret_cast<float4*>(go + 2 * H)[q] =
                make_float4(nn[q * 4], nn[q * 4 + 1], nn[q * 4 + 2], nn[q * 4 + 3]);
            reinterpret_cast<float4*>(go + 3 * H)[q] =
                make_float4(hn[q * 4], hn[q * 4 + 1], hn[q * 4 + 2], hn[q * 4 + 3]);
          }
        }
      }
      if (tid == 0) GRU_STAMP(10);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 64);
  }
}

// =============================================================================================
// backward (reverse of the forward time order of each direction)
// =============================================================================================
__global__ void __launch_bounds__(GRU_THREADS, 1)
gru_bwd_kernel(const __grid_constant__ CUtensorMap tm_d0, const __grid_constant__ CUtensorMap tm_d1,
               const GruBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int H = p.H, Bp = p.Bp, T = p.T;
  const int nC = H / GRU_HC;
  const int dir = blockIdx.x / nC;
  const int j0 = (blockIdx.x % nC) * GRU_HC;
  const int K3 = 3 * H;
  const int nchunks = (K3 + 63) / 64;
  constexpr int WCHUNK = 16 * 128;  // 16 rows x 64 bf16
  const int ring_bytes = p.ring * ring_stride(Bp) + (128 - Bp) * 128;
  const GruSmem s = carve(smem_raw, nchunks * WCHUNK, ring_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = p.ndir * H;
  const long long M = (long long)T * Bp;
  const CUtensorMap* tm = dir == 0 ? &tm_d0 : &tm_d1;

  for (int k = tid; k < ring_bytes / 16; k += GRU_THREADS)
    reinterpret_cast<uint4*>(s.ring)[k] = make_uint4(0, 0, 0, 0);
  {
    // resident operand: rows = the 16 hidden units k0..k0+15 of W_hh^T, K = 3H
    const int pieces_per_row = nchunks * 8;
    for (int k = tid; k < 16 * pieces_per_row; k += GRU_THREADS) {
      const int r = k / pieces_per_row, pc = k % pieces_per_row;
      const int col = pc * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (col < K3)
        v = *reinterpret_cast<const uint4*>(p.whhT + ((long long)dir * H + j0 + r) * K3 + col);
      *reinterpret_cast<uint4*>(s.wtile + (pc >> 3) * WCHUNK + sw128_offset(r, pc & 7)) = v;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < GRU_MAX_RING; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(s.accfull, 1);
    mbar_fence_init();
    tma_prefetch_desc(tm);
  }
  if (warp == 4) tmem_alloc(s.tmem_slot, 32);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s.tmem_slot;
  unsigned int* ctr = p.barrier + dir;

  if (warp == 5) {
    if (lane == 0) {
      unsigned int fill = 0;
      // the recurrent product is needed for every step except the last one processed
      for (int step = 0; step + 1 < T; ++step) {
        grid_wait(ctr, (unsigned int)nC * (step + 1));   // dgh of this step is complete
        fence_proxy_async_all();
        tma_gather(s, tm, (step & 1) * Bp, Bp, nchunks, p.ring, fill);
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      unsigned int fill = 0;
      for (int step = 0; step + 1 < T; ++step)
        mma_consume<16>(s, tmem_base, nchunks, WCHUNK, Bp, p.ring, fill);
    }
  } else {
    const bool active = tid < Bp;
    float dh_rec[GRU_HC];   // dL/dh_t arriving through the recurrence (own 16 units)
    float db_i[48], db_hn[GRU_HC];
#pragma unroll
    for (int jj = 0; jj < GRU_HC; ++jj) { dh_rec[jj] = 0.f; db_hn[jj] = 0.f; }
#pragma unroll
    for (int r = 0; r < 48; ++r) db_i[r] = 0.f;

    for (int step = 0; step < T; ++step) {
      // forward order of dir 0 is t=0..T-1, so its backward order is T-1..0; dir 1 mirrored
      const int t = dir == 0 ? (T - 1 - step) : step;
      const int tp = dir == 0 ? t - 1 : t + 1;       // time index of h_{prev} in forward order
      const bool has_prev = dir == 0 ? (t > 0) : (t < T - 1);
      bf16* xb = p.xchg + ((long long)(dir * 2 + (step & 1)) * Bp) * K3;
      const long long m = (long long)t * Bp + tid;
      // ---- saved activations of this step: independent of the recurrence, fetched first ----
      float rr[GRU_HC], zz[GRU_HC], nn[GRU_HC], hn[GRU_HC], dh[GRU_HC], hp[GRU_HC];
      if (active) {
        const float* go = p.gates + ((m * p.ndir + dir) * 4) * H + j0;
        const float* dyo = p.dy + m * D + dir * H + j0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(go) + q);
          const float4 b = __ldg(reinterpret_cast<const float4*>(go + H) + q);
          const float4 c = __ldg(reinterpret_cast<const float4*>(go + 2 * H) + q);
          const float4 d = __ldg(reinterpret_cast<const float4*>(go + 3 * H) + q);
          const float4 e = __ldg(reinterpret_cast<const float4*>(dyo) + q);
          rr[q * 4] = a.x; rr[q * 4 + 1] = a.y; rr[q * 4 + 2] = a.z; rr[q * 4 + 3] = a.w;
          zz[q * 4] = b.x; zz[q * 4 + 1] = b.y; zz[q * 4 + 2] = b.z; zz[q * 4 + 3] = b.w;
          nn[q * 4] = c.x; nn[q * 4 + 1] = c.y; nn[q * 4 + 2] = c.z; nn[q * 4 + 3] = c.w;
          hn[q * 4] = d.x; hn[q * 4 + 1] = d.y; hn[q * 4 + 2] = d.z; hn[q * 4 + 3] = d.w;
          dh[q * 4] = e.x; dh[q * 4 + 1] = e.y; dh[q * 4 + 2] = e.z; dh[q * 4 + 3] = e.w;
        }
        if (has_prev) {
          const float* yo = p.y + ((long long)tp * Bp + tid) * D + dir * H + j0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(yo) + q);
            hp[q * 4] = a.x; hp[q * 4 + 1] = a.y; hp[q * 4 + 2] = a.z; hp[q * 4 + 3] = a.w;
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < GRU_HC; ++jj) hp[jj] = 0.f;
        }
      }
      // ---- recurrent part of dL/dh_t: product issued in the previous step ----
      if (step > 0) {
        mbar_wait(s.accfull, (step - 1) & 1);
        tc_fence_after_sync();
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
        tmem_ld_wait();
        tc_fence_before_sync();
#pragma unroll
        for (int jj = 0; jj < GRU_HC; ++jj) dh_rec[jj] += __uint_as_float(v[jj]);
      }
      float dr[GRU_HC], dz[GRU_HC], dn[GRU_HC], dnr[GRU_HC];
      if (active) {
#pragma unroll
        for (int jj = 0; jj < GRU_HC; ++jj) {
          const float g = dh[jj] + dh_rec[jj];
          dn[jj] = g * (1.f - zz[jj]) * (1.f - nn[jj] * nn[jj]);
          dz[jj] = g * (hp[jj] - nn[jj]) * zz[jj] * (1.f - zz[jj]);
          dr[jj] = dn[jj] * hn[jj] * rr[jj] * (1.f - rr[jj]);
          dnr[jj] = dn[jj] * rr[jj];
          dh_rec[jj] = g * zz[jj];      // direct path h_{t-1} -> h_t; the W_hh path is added next step
          db_i[jj] += dr[jj];
          db_i[16 + jj] += dz[jj];
          db_i[32 + jj] += dn[jj];
          db_hn[jj] += dnr[jj];
        }
        // critical path: the exchange rows [dr | dz | dn*r] every CTA gathers for the product
        if (step + 1 < T) {
          uint4 pk[2];
          uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
          bf16* x = xb + (long long)tid * K3 + j0;
#pragma unroll
          for (int q = 0; q < 8; ++q) pw[q] = pack_bf16x2(dr[2 * q], dr[2 * q + 1]);
          reinterpret_cast<uint4*>(x)[0] = pk[0]; reinterpret_cast<uint4*>(x)[1] = pk[1];
#pragma unroll
          for (int q = 0; q < 8; ++q) pw[q] = pack_bf16x2(dz[2 * q], dz[2 * q + 1]);
          reinterpret_cast<uint4*>(x + H)[0] = pk[0]; reinterpret_cast<uint4*>(x + H)[1] = pk[1];
#pragma unroll
          for (int q = 0; q < 8; ++q) pw[q] = pack_bf16x2(dnr[2 * q], dnr[2 * q + 1]);
          reinterpret_cast<uint4*>(x + 2 * H)[0] = pk[0];
          reinterpret_cast<uint4*>(x + 2 * H)[1] = pk[1];
          fence_proxy_async_all();
        }
      }
      if (step + 1 < T) {
        epi_barrier();
        if (tid == 0) {
          __threadfence();
          grid_arrive(ctr);
        }
      }
      // ---- off the critical path: operands of the dX / dW GEMMs ----
      if (active) {
        bf16* o = p.dgi + m * (p.ndir * K3) + dir * K3 + j0;
        uint4 pk[2];
        uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
        for (int q = 0; q < 8; ++q) pw[q] = pack_bf16x2(dr[2 * q], dr[2 * q + 1]);
        reinterpret_cast<uint4*>(o)[0] = pk[0]; reinterpret_cast<uint4*>(o)[1] = pk[1];
#pragma unroll
        for (int q = 0; q < 8; ++q) pw[q] = pack_bf16x2(dz[2 * q], dz[2 * q + 1]);
        reinterpret_cast<uint4*>(o + H)[0] = pk[0]; reinterpret_cast<uint4*>(o + H)[1] = pk[1];
#pragma unroll
        for (int q = 0; q < 8; ++q) pw[q] = pack_bf16x2(dn[2 * q], dn[2 * q + 1]);
        reinterpret_cast<uint4*>(o + 2 * H)[0] = pk[0];
        reinterpret_cast<uint4*>(o + 2 * H)[1] = pk[1];
        bf16* gt = p.dgiT + ((long long)dir * K3 + j0) * M + m;
        bf16* nt = p.dghnT + ((long long)dir * H + j0) * M + m;
#pragma unroll
        for (int jj = 0; jj < GRU_HC; ++jj) {
          gt[(long long)jj * M] = __float2bfloat16_rn(dr[jj]);
          gt[(long long)(H + jj) * M] = __float2bfloat16_rn(dz[jj]);
          gt[(long long)(2 * H + jj) * M] = __float2bfloat16_rn(dn[jj]);
          nt[(long long)jj * M] = __float2bfloat16_rn(dnr[jj]);
        }
      }
    }
    // ---- bias gradients: reduce the per-batch-row partial sums over the CTA ----
    epi_barrier();
    float* red = s.scratch;  // 64 floats
    if (tid < 64) red[tid] = 0.f;
    epi_barrier();
#pragma unroll
    for (int r = 0; r < 48; ++r) {
      const float v = warp_sum(active ? db_i[r] : 0.f);
      if (lane == 0) atomicAdd(&red[r], v);
    }
#pragma unroll
    for (int jj = 0; jj < GRU_HC; ++jj) {
      const float v = warp_sum(active ? db_hn[jj] : 0.f);
      if (lane == 0) atomicAdd(&red[48 + jj], v);
    }
    epi_barrier();
    if (tid < 48) {
      const int g = tid / GRU_HC, jj = tid % GRU_HC;
      const int idx = dir * K3 + g * H + j0 + jj;
      atomicAdd(p.dbih + idx, red[tid]);
      atomicAdd(p.dbhh + idx, g < 2 ? red[tid] : red[48 + jj]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 32);
  }
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols,
                      long long ld, int box_rows);

// ring slots that fit next to the resident weights
static int gru_ring_slots(int wbytes, int Bp, int nchunks, size_t* smem_bytes) {
  const int stride = Bp * 128;
  const int slack = (128 - Bp) * 128;
  const int fixed = wbytes + slack + 1024 /*align*/ + (2 * GRU_MAX_RING + 2) * 8 + 256 /*scratch*/;
  int ring = (227 * 1024 - fixed) / stride;
  if (ring > GRU_MAX_RING) ring = GRU_MAX_RING;
  if (ring > nchunks) ring = nchunks;
  if (ring < 2) return -1;
  *smem_bytes = (size_t)fixed + (size_t)ring * stride;
  return ring;
}

static int gru_launch(const void* kernel, int grid, size_t smem, void** args, cudaStream_t stream) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return SB_ERR_CUDA;
  if (cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(GRU_THREADS), args, smem, stream) !=
      cudaSuccess)
    return SB_ERR_CUDA;
  return SB_OK;
}

}  // namespace sb

using namespace sb;

static unsigned long long* g_gru_dbg = nullptr;

// developer hook: device buffer of >= 64*16 u64 that receives a per-step timeline of CTA 0 of the
// next sb_gru_fwd launches (nullptr disables).  Not part of the drop-in surface.
extern "C" int sb_debug_gru_timeline(void* dev_buffer) {
  g_gru_dbg = reinterpret_cast<unsigned long long*>(dev_buffer);
  return SB_OK;
}

static int gru_check(int T, int Bp, int H, int ndir) {
  if (T <= 0 || Bp <= 0 || H <= 0 || (ndir != 1 && ndir != 2)) return SB_ERR_INVALID;
  if (H % GRU_HC != 0 || Bp % 8 != 0 || Bp > 128) return SB_ERR_UNSUPPORTED;
  if (ndir * (H / GRU_HC) > sb::device_sm_count()) return SB_ERR_UNSUPPORTED;
  return SB_OK;
}

extern "C" int sb_gru_fwd(const float* gi, const void* whh_bf16, const float* bhh, float* y,
                          void* xn_bf16, void* xnT_bf16, float* gates, unsigned int* barrier,
                          int T, int Bp, int H, int ndir, void* stream_) {
  int rc = gru_check(T, Bp, H, ndir);
  if (rc != SB_OK) return rc;
  if (!gi || !whh_bf16 || !bhh || !y || !xn_bf16 || !barrier) return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GruFwdParams p;
  p.gi = gi; p.whh = reinterpret_cast<const bf16*>(whh_bf16); p.bhh = bhh; p.y = y;
  p.xn = reinterpret_cast<bf16*>(xn_bf16); p.xnT = reinterpret_cast<bf16*>(xnT_bf16);
  p.gates = gates; p.barrier = barrier; p.T = T; p.Bp = Bp; p.H = H; p.ndir = ndir;
  p.dbg = g_gru_dbg;
  const int nchunks = (H + 63) / 64;
  size_t smem = 0;
  p.ring = gru_ring_slots(nchunks * 48 * 128, Bp, nchunks, &smem);
  if (p.ring < 0) return SB_ERR_UNSUPPORTED;
  // one tensor map per direction over that direction's H columns of h (bf16 [T*Bp][ndir*H]):
  // columns past H are out of bounds and read as zero
  CUtensorMap tm[2];
  for (int d = 0; d < 2; ++d) {
    const int dd = d < ndir ? d : 0;
    rc = make_tmap_bf16_2d(&tm[d], p.xn + (size_t)dd * H, (long long)T * Bp, H,
                           (long long)ndir * H, Bp);
    if (rc != SB_OK) return rc;
  }
  if (cudaMemsetAsync(barrier, 0, sizeof(unsigned int) * ndir, stream) != cudaSuccess)
    return SB_ERR_CUDA;
  void* args[] = {(void*)&tm[0], (void*)&tm[1], (void*)&p};
  return gru_launch((const void*)gru_fwd_kernel, ndir * (H / GRU_HC), smem, args, stream);
}

extern "C" int sb_gru_bwd_workspace_size(int Bp, int H, int ndir, size_t* bytes) {
  if (!bytes || Bp <= 0 || H <= 0) return SB_ERR_INVALID;
  *bytes = (size_t)ndir * 2 * Bp * 3 * H * sizeof(bf16) + 256;
  return SB_OK;
}

extern "C" int sb_gru_bwd(const float* dy, const float* y, const float* gates,
                          const void* whhT_bf16, void* dgi_bf16, void* dgiT_bf16,
                          void* dghnT_bf16, float* dbih, float* dbhh, void* workspace,
                          size_t workspace_bytes, unsigned int* barrier, int T, int Bp, int H,
                          int ndir, void* stream_) {
  int rc = gru_check(T, Bp, H, ndir);
  if (rc != SB_OK) return rc;
  if (!dy || !y || !gates || !whhT_bf16 || !dgi_bf16 || !dgiT_bf16 || !dghnT_bf16 || !dbih ||
      !dbhh || !workspace || !barrier)
    return SB_ERR_INVALID;
  size_t need = 0;
  sb_gru_bwd_workspace_size(Bp, H, ndir, &need);
  if (workspace_bytes < need) return SB_ERR_WORKSPACE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GruBwdParams p;
  p.dy = dy; p.y = y; p.gates = gates; p.whhT = reinterpret_cast<const bf16*>(whhT_bf16);
  p.dgi = reinterpret_cast<bf16*>(dgi_bf16); p.dgiT = reinterpret_cast<bf16*>(dgiT_bf16);
  p.dghnT = reinterpret_cast<bf16*>(dghnT_bf16);
  p.xchg = reinterpret_cast<bf16*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  p.dbih = dbih; p.dbhh = dbhh; p.barrier = barrier; p.T = T; p.Bp = Bp; p.H = H; p.ndir = ndir;
  const int K3 = 3 * H;
  const int nchunks = (K3 + 63) / 64;
  size_t smem = 0;
  p.ring = gru_ring_slots(nchunks * 16 * 128, Bp, nchunks, &smem);
  if (p.ring < 0) return SB_ERR_UNSUPPORTED;
  // per direction: the two parity buffers stacked as [2*Bp rows][3H cols]
  CUtensorMap tm[2];
  for (int d = 0; d < 2; ++d) {
    const int dd = d < ndir ? d : 0;
    rc = make_tmap_bf16_2d(&tm[d], p.xchg + (size_t)dd * 2 * Bp * K3, 2LL * Bp, K3, K3, Bp);
    if (rc != SB_OK) return rc;
  }
  if (cudaMemsetAsync(barrier, 0, sizeof(unsigned int) * ndir, stream) != cudaSuccess)
    return SB_ERR_CUDA;
  void* args[] = {(void*)&tm[0], (void*)&tm[1], (void*)&p};
  return gru_launch((const void*)gru_bwd_kernel, ndir * (H / GRU_HC), smem, args, stream);
}
