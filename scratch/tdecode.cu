// sb_rnnt_decode_static: transducer beam search over a PRECOMPUTED (teacher-forced) lattice, on the
// GPU, one CTA per utterance.
//
// Replaces `transducer.decoders.decode_static(lp, beam_size, blank)` of the un-vendored
// awni/transducer (imported at speech/models/transducer_model.py:10, called per utterance on a
// host numpy array at :92-101).  The dependency's source is not available, so the algorithm is
// the standard transducer beam search (Graves 2012, section 3) restricted to a static lattice:
// because lp[t, u, :] was computed with teacher forcing, the prediction-network state of a
// hypothesis is the number u of labels it has emitted.  It is restated in oracle/
// decode_static_ref.py (the CPU oracle, pinned by exhaustive enumeration on tiny lattices); this
// kernel reproduces that restatement step for step, including its tie order:
//   per frame t:  frontier = beam; done = {}
//     repeat: every frontier hypothesis (in order) adds score + lp[t,u,blank] to done[hyp]
//             (log-sum-exp when the hypothesis is already there: prefix merging);
//             every (hypothesis i, label k != blank) with u+1 < U is a candidate
//             score + lp[t,u,k]; the next frontier = the beam_size best candidates by
//             (score desc, i*V + k asc) = a stable descending sort of the candidate list
//     beam = the beam_size best of done by (score desc, insertion order asc)
//   result = beam[0] after the last frame.
// Hypotheses are nodes (parent, label) of a per-utterance trie made canonical by a hash map, so
// "the same label sequence" is an integer comparison.  Scores are float64, as Python floats.
// Roofline: HBM-nominal (the lattice is read once); latency-bound by the T x U serial expansion.
#include "common.cuh"
#include <math.h>

#include "../../include/speech_b200.h"

namespace sb {

static constexpr int TD_THREADS = 128;
static constexpr int TD_MAX_BEAM = 32;

struct TdParams {
  const float* lp;      // (B, T, U1, V) log-probabilities
  const int* tlens;     // (B) frames to search (<= T)
  const int* ulens;     // (B) lattice rows U = labels + 1 (<= U1)
  int* nodes;           // (B, 2*node_cap) trie arena: parent, label
  unsigned long long* hkeys;   // (B, hcap)
  int* hvals;                  // (B, hcap)
  double* done_score;   // (B, done_cap)
  int* done_node;       // (B, done_cap)
  int* done_len;        // (B, done_cap)
  int* out_labels;      // (B, U1)
  int* out_lens;        // (B)
  double* out_scores;   // (B) log-probability of the best hypothesis
  int B, T, U1, V, K, blank, hcap, node_cap, done_cap;
};

SB_DEVINL double td_lse(double a, double b) {
  const double m = fmax(a, b);
  return m + log(exp(a - m) + exp(b - m));
}

// block arg-max over (score desc, index asc) of the live entries of sc[0..n); returns the index
// (or -1) to every thread.  `taken` entries carry NaN.
SB_DEVINL int td_argmax(const double* sc, int n, double* red_s, int* red_i, int tid) {
  double bs = 0.0;
  int bi = -1;
  for (int i = tid; i < n; i += TD_THREADS) {
    const double v = sc[i];
    if (isnan(v)) continue;
    if (bi < 0 || v > bs) { bs = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
  }
  if ((tid & 31) == 0) { red_s[tid >> 5] = bs; red_i[tid >> 5] = bi; }
  __syncthreads();
  bs = red_s[0];
  bi = red_i[0];
  for (int w = 1; w < TD_THREADS / 32; ++w)
    if (red_i[w] >= 0 && (bi < 0 || red_s[w] > bs || (red_s[w] == bs && red_i[w] < bi))) {
      bs = red_s[w];
      bi = red_i[w];
    }
  __syncthreads();
  return bi;
}

__global__ void __launch_bounds__(TD_THREADS) rnnt_decode_static_kernel(const TdParams p) {
  extern __shared__ unsigned char td_smem[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int K = p.K, V = p.V;
  const int T = min(p.tlens[b], p.T), U = min(p.ulens[b], p.U1);
  double* cand = reinterpret_cast<double*>(td_smem);       // [K * V]
  __shared__ double beam_s[TD_MAX_BEAM], fr_s[TD_MAX_BEAM], new_s[TD_MAX_BEAM];
  __shared__ int beam_n[TD_MAX_BEAM], beam_l[TD_MAX_BEAM];
  __shared__ int fr_n[TD_MAX_BEAM], fr_l[TD_MAX_BEAM];
  __shared__ int new_n[TD_MAX_BEAM], new_l[TD_MAX_BEAM];
  __shared__ double red_s[TD_THREADS / 32];
  __shared__ int red_i[TD_THREADS / 32];
  __shared__ int nbeam, nfr, ndone, found, round_id;

  int* nodes = p.nodes + (size_t)b * 2 * p.node_cap;
  unsigned long long* hkeys = p.hkeys + (size_t)b * p.hcap;
  int* hvals = p.hvals + (size_t)b * p.hcap;
  double* dscore = p.done_score + (size_t)b * p.done_cap;
  int* dnode = p.done_node + (size_t)b * p.done_cap;
  int* dlen = p.done_len + (size_t)b * p.done_cap;
  const float* lp = p.lp + (size_t)b * p.T * p.U1 * V;

  for (int k = tid; k < p.hcap; k += TD_THREADS) hkeys[k] = 0ull;
  if (tid == 0) {
    nbeam = 1; beam_s[0] = 0.0; beam_n[0] = 0; beam_l[0] = 0;
    nodes[0] = -1; nodes[1] = -1;
    round_id = 0;
  }
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    if (tid < nbeam) { fr_s[tid] = beam_s[tid]; fr_n[tid] = beam_n[tid]; fr_l[tid] = beam_l[tid]; }
    if (tid == 0) { nfr = nbeam; ndone = 0; }
    __syncthreads();
    for (int level = 0; level < U; ++level) {
      // ---- 1. blank emissions into `done`, hypothesis by hypothesis (merging order matters) ----
      const int nf = nfr;
      for (int i = 0; i < nf; ++i) {
        const int u = fr_l[i];
        if (u >= U) continue;                       // (uniform: shared data)
        if (tid == 0) found = -1;
        __syncthreads();
        const int nd = ndone, me = fr_n[i];
        for (int j = tid; j < nd; j += TD_THREADS)
          if (dnode[j] == me) found = j;            // at most one entry holds this hypothesis
        __syncthreads();
        if (tid == 0) {
          const double bsc = fr_s[i] + (double)lp[((size_t)t * p.U1 + u) * V + p.blank];
          if (found >= 0) dscore[found] = td_lse(dscore[found], bsc);
          else if (nd < p.done_cap) { dscore[nd] = bsc; dnode[nd] = me; dlen[nd] = u; ndone = nd + 1; }
        }
        __syncthreads();
      }
      // ---- 2. label expansions: candidates (i, k), the K best form the next frontier ----
      bool any = false;
      for (int c = tid; c < K * V; c += TD_THREADS) {
        const int i = c / V, k = c - i * V;
        double sc = nan("");
        if (i < nf && k != p.blank && fr_l[i] + 1 < U) {
          sc = fr_s[i] + (double)lp[((size_t)t * p.U1 + fr_l[i]) * V + k];
          any = true;
        }
        cand[c] = sc;
      }
      const int have = __syncthreads_or(any ? 1 : 0);
      if (!have) break;
      int nsel = 0;
      for (int r = 0; r < K; ++r) {
        const int c = td_argmax(cand, K * V, red_s, red_i, tid);
        if (c < 0) break;
        if (tid == 0) {
          const int i = c / V, k = c - i * V;
          // canonical node of hypothesis i extended by label k
          const unsigned long long key =
              ((unsigned long long)(unsigned int)fr_n[i] << 32) | (unsigned int)(k + 1);
          unsigned int h = (unsigned int)((key * 0x9E3779B97F4A7C15ull) >> 32) & (p.hcap - 1);
          int id = -1;
          for (int probe = 0; probe < p.hcap; ++probe) {
            const unsigned long long kk = hkeys[h];
            if (kk == key) { id = hvals[h]; break; }
            if (kk == 0ull) {
              id = 1 + round_id * K + nsel;
              if (id >= p.node_cap) { id = -1; break; }
              hkeys[h] = key;
              hvals[h] = id;
              nodes[2 * id] = fr_n[i];
              nodes[2 * id + 1] = k;
              break;
            }
            h = (h + 1) & (p.hcap - 1);
          }
          new_s[nsel] = cand[c];
          new_n[nsel] = id;
          new_l[nsel] = fr_l[i] + 1;
          cand[c] = nan("");
        }
        ++nsel;
        __syncthreads();
      }
      if (tid < nsel) { fr_s[tid] = new_s[tid]; fr_n[tid] = new_n[tid]; fr_l[tid] = new_l[tid]; }
      if (tid == 0) { nfr = nsel; ++round_id; }
      __syncthreads();
    }
    // ---- 3. beam = the K best of `done` ----
    const int nd = ndone;
    int nb = 0;
    for (int r = 0; r < K; ++r) {
      const int j = td_argmax(dscore, nd, red_s, red_i, tid);
      if (j < 0) break;
      if (tid == 0) {
        beam_s[nb] = dscore[j]; beam_n[nb] = dnode[j]; beam_l[nb] = dlen[j];
        dscore[j] = nan("");
      }
      ++nb;
      __syncthreads();
    }
    if (tid == 0) nbeam = nb;
    __syncthreads();
  }

  if (tid == 0) {
    int len = 0;
    const int n = nbeam > 0 ? beam_n[0] : 0;
    for (int q = n; q > 0; q = nodes[2 * q]) ++len;
    int* out = p.out_labels + (size_t)b * p.U1;
    int k = len;
    for (int q = n; q > 0; q = nodes[2 * q]) out[--k] = nodes[2 * q + 1];
    p.out_lens[b] = len;
    p.out_scores[b] = nbeam > 0 ? beam_s[0] : -INFINITY;
  }
}

static int td_pow2(size_t need) {
  size_t cap = 64;
  while (cap < need) cap <<= 1;
  return (int)cap;
}
struct TdSizes { size_t node_cap, hcap, done_cap, total; };
static TdSizes td_sizes(int B, int T, int U1, int K) {
  TdSizes s;
  s.node_cap = (size_t)T * U1 * K + 2;
  s.hcap = (size_t)td_pow2(2 * s.node_cap);
  s.done_cap = (size_t)K * (U1 + 1) + 1;
  s.total = 1024 + (size_t)B * (2 * s.node_cap * sizeof(int) +
                                s.hcap * (sizeof(unsigned long long) + sizeof(int)) +
                                s.done_cap * (sizeof(double) + 2 * sizeof(int))) + 4096;
  return s;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_rnnt_decode_static_workspace_size(int B, int T, int U1, int beam_size,
                                                    size_t* bytes) {
  if (!bytes || B <= 0 || T <= 0 || U1 <= 0 || beam_size <= 0) return SB_ERR_INVALID;
  *bytes = td_sizes(B, T, U1, beam_size).total;
  return SB_OK;
}

extern "C" int sb_rnnt_decode_static(const float* lp, const int* tlens, const int* ulens, int B,
                                     int T, int U1, int V, int beam_size, int blank,
                                     int* out_labels, int* out_lens, double* out_scores,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  if (!lp || !tlens || !ulens || !out_labels || !out_lens || !out_scores || !workspace)
    return SB_ERR_INVALID;
  if (B <= 0 || T <= 0 || U1 <= 0 || V <= 0 || blank < 0 || blank >= V || beam_size <= 0)
    return SB_ERR_INVALID;
  if (beam_size > TD_MAX_BEAM) return SB_ERR_UNSUPPORTED;
  const TdSizes sz = td_sizes(B, T, U1, beam_size);
  if (workspace_bytes < sz.total) return SB_ERR_WORKSPACE;
  const size_t smem = (size_t)beam_size * V * sizeof(double) + 16;
  if (smem > 200 * 1024) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute(rnnt_decode_static_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem) != cudaSuccess)
    return SB_ERR_CUDA;
  TdParams p;
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  p.hkeys = reinterpret_cast<unsigned long long*>(w);
  w += (size_t)B * sz.hcap * sizeof(unsigned long long);
  p.done_score = reinterpret_cast<double*>(w);
  w += (size_t)B * sz.done_cap * sizeof(double);
  p.hvals = reinterpret_cast<int*>(w);
  w += (size_t)B * sz.hcap * sizeof(int);
  p.nodes = reinterpret_cast<int*>(w);
  w += (size_t)B * 2 * sz.node_cap * sizeof(int);
  p.done_node = reinterpret_cast<int*>(w);
  w += (size_t)B * sz.done_cap * sizeof(int);
  p.done_len = reinterpret_cast<int*>(w);
  p.lp = lp; p.tlens = tlens; p.ulens = ulens;
  p.out_labels = out_labels; p.out_lens = out_lens; p.out_scores = out_scores;
  p.B = B; p.T = T; p.U1 = U1; p.V = V; p.K = beam_size; p.blank = blank;
  p.hcap = (int)sz.hcap; p.node_cap = (int)sz.node_cap; p.done_cap = (int)sz.done_cap;
  rnnt_decode_static_kernel<<<B, TD_THREADS, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
