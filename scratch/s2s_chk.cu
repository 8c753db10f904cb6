// Attention decoder step of the sequence-to-sequence model: forward and backward kernels.
//
// Replaces, for Seq2Seq.decode / decode_step (speech/models/seq2seq.py:78-137) and
// NNAttention.forward (:344-360), the per-token chain of small library calls the reference runs
// (nn.Embedding, nn.GRUCell -> cuBLAS, Conv1d -> cuDNN, broadcast add / ReLU / Linear / softmax /
// weighted sum -> ATen, LinearND -> cuBLAS; ~8 launches and a (B,T,H) temporary per token):
//   step u:  ix = emb[y_u] + sx_{u-1}                                   (:84,:100-101)
//            hx_u = GRUCell(ix, hx_{u-1})                               (:103)
//            score_t = w . relu(eh_t + hx_u + conv1d(ax_{u-1})_t) + b   (:345-353)
//            ax_u = softmax_t(score [* log T]);  sx_u = sum_t ax_u[t] eh_t   (:354-359)
//            out_u = fc(hx_u + sx_u)                                    (:108)
// as TWO kernels per token in each direction of time:
//   s2s_cell_fwd      embedding gather + context add + GRU cell, fp32 on CUDA cores (the (B x 2H)
//                     x (2H x 3H) product of one token is 0.1 GFLOP: launch-latency-, not
//                     throughput-bound; one warp per hidden unit, lanes over the batch);
//   s2s_attn_fwd      one CTA per utterance: ONE pass over the encoder states with an online
//                     softmax (the reference reads them twice and materialises a (B,T,H)
//                     temporary), then the output projection, and for the decode path the
//                     arg-max token / log-softmax of the step, so that greedy and beam decoding
//                     never leave the device;
//   s2s_attn_bwd      gradient of the step's attention + output projection for one utterance:
//                     two passes over the encoder states (softmax Jacobian needs sum_t a_t da_t),
//                     accumulates d eh in place, emits d ax_{u-1}, the per-utterance parameter
//                     gradients of the attention, and the gate pre-activation gradients of the
//                     cell (so the cell's backward is a pure matrix product);
//   s2s_cell_bwd      d ix = d gi W_ih, d hx_{u-1} = d gh W_hh + z * d hx_u (lanes over the batch).
// The weight gradients of the cell, the embedding and fc are time-batched contractions over all
// (u, b) rows and run once per sequence on the tcgen05 GEMM (SB_GEMM_A_MN | SB_GEMM_B_MN).
// Everything is fp32 (the reference's arithmetic); roofline: HBM/L2 bandwidth on eh per token
// (B*T*H*4 bytes forward, 3x that backward), in practice launch/latency-bound at B <= 64.
#include "../speech_b200/csrc/common.cuh"
#include <math.h>
#include <string.h>

#include "../include/speech_b200.h"

namespace sb {

static constexpr int S2S_UPC = 8;        // hidden units per CTA of the cell kernels (one per warp)
static constexpr int S2S_KC = 512;       // K chunk staged in shared memory
static constexpr int ATT_THREADS = 256;
static constexpr int ATT_MAXR = 32;      // H <= 32*32 = 1024

// ------------------------------------------------------------------------------------------------
// cell forward
// ------------------------------------------------------------------------------------------------
struct CellFwdParams {
  const float* emb;       // [Vocab][H]
  const int* tok;         // token of row b at tok[b * tok_stride]
  int tok_stride;
  const float* sx_prev;   // [B][H] or null (first step)
  const float* hx_prev;   // [B][H]
  const float* w_ih;      // [3H][H]
  const float* w_hh;      // [3H][H]
  const float* b_ih;      // [3H]
  const float* b_hh;      // [3H]
  float* hx;              // [B][H]
  float* ix_save;         // [B][H] or null
  float* gates_save;      // [B][4][H] (r, z, n, hn) or null
  const int* done;        // device flag: != 0 -> the decode has finished, do nothing (or null)
  int B, H;
};

__global__ void __launch_bounds__(32 * S2S_UPC) s2s_cell_fwd_kernel(const CellFwdParams p) {
  extern __shared__ float cell_smem[];
  if (p.done && *p.done) return;
  const int H = p.H, B = p.B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int j = blockIdx.x * S2S_UPC + warp;           // this warp's hidden unit
  const int ld = S2S_KC + 4;                            // padded row: conflict-free float4 reads
  float* xs = cell_smem;                                // [32][ld]  ix chunk
  float* hs = cell_smem + 32 * ld;                      // [32][ld]  hx_prev chunk
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int b = b0 + lane;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // gi_r, gi_z, gi_n, gh_r, gh_z, gh_n
    for (int k0 = 0; k0 < H; k0 += S2S_KC) {
      const int kc = min(S2S_KC, H - k0);
      __syncthreads();
      // stage ix = emb[tok] + sx_prev and hx_prev for rows b0..b0+31, columns k0..k0+kc
      for (int e = tid; e < 32 * (kc / 4); e += 32 * S2S_UPC) {
        const int r = e / (kc / 4), c4 = e % (kc / 4);
        const int rb = b0 + r;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), h = x;
        if (rb < B) {
          const int tk = p.tok[(long long)rb * p.tok_stride];
          x = __ldg(reinterpret_cast<const float4*>(p.emb + (long long)tk * H + k0) + c4);
          if (p.sx_prev) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(p.sx_prev + (long long)rb * H + k0) + c4);
            x.x += s.x; x.y += s.y; x.z += s.z; x.w += s.w;
          }
          h = __ldg(reinterpret_cast<const float4*>(p.hx_prev + (long long)rb * H + k0) + c4);
          if (p.ix_save && blockIdx.x == 0)
            reinterpret_cast<float4*>(p.ix_save + (long long)rb * H + k0)[c4] = x;
        }
        reinterpret_cast<float4*>(xs + r * ld)[c4] = x;
        reinterpret_cast<float4*>(hs + r * ld)[c4] = h;
      }
      __syncthreads();
      if (j < H) {
        const float* wi = p.w_ih + (long long)j * H + k0;
        const float* wh = p.w_hh + (long long)j * H + k0;
        const float4* xr = reinterpret_cast<const float4*>(xs + lane * ld);
        const float4* hr = reinterpret_cast<const float4*>(hs + lane * ld);
#pragma unroll 4
        for (int c4 = 0; c4 < kc / 4; ++c4) {
          const float4 x = xr[c4], h = hr[c4];
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(wi + (long long)g * H * H) + c4);
            const float4 c = __ldg(reinterpret_cast<const float4*>(wh + (long long)g * H * H) + c4);
            acc[g] += a.x * x.x + a.y * x.y + a.z * x.z + a.w * x.w;
            acc[3 + g] += c.x * h.x + c.y * h.y + c.z * h.z + c.w * h.w;
          }
        }
      }
    }
    if (j < H && b < B) {
      const float gir = acc[0] + p.b_ih[j], giz = acc[1] + p.b_ih[H + j], gin = acc[2] + p.b_ih[2 * H + j];
      const float ghr = acc[3] + p.b_hh[j], ghz = acc[4] + p.b_hh[H + j], ghn = acc[5] + p.b_hh[2 * H + j];
      const float r = 1.f / (1.f + expf(-(gir + ghr)));
      const float z = 1.f / (1.f + expf(-(giz + ghz)));
      const float n = tanhf(gin + r * ghn);
      const float hp = p.hx_prev[(long long)b * H + j];
      p.hx[(long long)b * H + j] = (1.f - z) * n + z * hp;
      if (p.gates_save) {
        float* g = p.gates_save + (long long)b * 4 * H + j;
        g[0] = r; g[H] = z; g[2 * H] = n; g[3 * H] = ghn;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention + output projection forward (one CTA per utterance / beam entry)
// ------------------------------------------------------------------------------------------------
struct AttnFwdParams {
  const float* eh;        // (Beh, T, H) encoder states; row b uses utterance b % Beh... see eh_bcast
  int eh_bcast;           // 1: every row attends over utterance 0 (beam search of one utterance)
  const float* hx;        // (B, H) decoder state of this step
  const float* ax_prev;   // (B, T) or null
  const float* conv_w;    // (H, Kc)
  const float* conv_b;    // (H)
  const float* lin_w;     // (H)
  float lin_b;
  float* sx;              // (B, H)
  float* ax;              // (B, T)
  // output projection (optional): logits[b*logit_stride + c] = fc_b[c] + fc_w[c,:] . (hx + sx)
  const float* fc_w;      // (C, H) or null
  const float* fc_b;      // (C)
  float* logits;
  long long logit_stride;
  float* logp;            // (B, C) log-softmax of the logits or null
  int* argmax;            // (B) arg-max class (first maximum) or null
  int* history;           // greedy decode: history[b * hist_stride + hist_col] = arg-max (or null)
  int hist_stride, hist_col;
  int* end_count;         // += 1 when this row's arg-max == end_tok (or null)
  int end_tok;
  const int* done;        // device flag: != 0 -> do nothing
  int B, T, H, Kc, C, log_t;
};

__global__ void __launch_bounds__(ATT_THREADS) s2s_attn_fwd_kernel(const AttnFwdParams p) {
  extern __shared__ float att_smem[];
  if (p.done && *p.done) return;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, H = p.H, Kc = p.Kc, pad = (Kc - 1) / 2;
  float* axp = att_smem;                       // [T + Kc - 1]  zero-padded previous alignment
  float* cw = axp + T + Kc - 1;                // [H * Kc]
  float* dc = cw + H * Kc;                     // [H] hx + conv bias
  float* lw = dc + H;                          // [H]
  float* score = lw + H;                       // [T]
  float* wstat = score + T;                    // [8][2] per-warp (max, sum)
  float* wsx = wstat + 16;                     // [8][H] per-warp weighted sums; later o = hx + sx
  const bool has_prev = p.ax_prev != nullptr;

  for (int k = tid; k < T + Kc - 1; k += ATT_THREADS) {
    const int t = k - pad;
    axp[k] = (has_prev && t >= 0 && t < T) ? p.ax_prev[(size_t)b * T + t] : 0.f;
  }
  if (has_prev)
    for (int k = tid; k < H * Kc; k += ATT_THREADS) cw[k] = p.conv_w[k];
  for (int h = tid; h < H; h += ATT_THREADS) {
    dc[h] = p.hx[(size_t)b * H + h] + (has_prev ? p.conv_b[h] : 0.f);
    lw[h] = p.lin_w[h];
  }
  __syncthreads();

  const float tscale = p.log_t ? logf((float)T) : 1.0f;
  const float* eh = p.eh + (p.eh_bcast ? 0 : (size_t)b * T * H);
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
  float sxv[(ATT_MAXR * 32 + ATT_THREADS - 1) / ATT_THREADS];
  {
    int q = 0;
    for (int h = tid; h < H; h += ATT_THREADS, ++q) {
      float s = 0.f;
      for (int w = 0; w < ATT_THREADS / 32; ++w)
        if (wstat[w * 2] != -INFINITY) s += wsx[w * H + h] * __expf(wstat[w * 2] - m);
      sxv[q] = s * inv;
      p.sx[(size_t)b * H + h] = sxv[q];
    }
  }
  for (int t = tid; t < T; t += ATT_THREADS) p.ax[(size_t)b * T + t] = __expf(score[t] - m) * inv;
  if (!p.fc_w) return;
  // ---- output projection on o = hx + sx (seq2seq.py:108,131-132) ----
  __syncthreads();                 // everyone is done reading wsx
  float* o = wsx;                  // [H]
  float* lg = wsx + H;             // [C]
  {
    int q = 0;
    for (int h = tid; h < H; h += ATT_THREADS, ++q) o[h] = p.hx[(size_t)b * H + h] + sxv[q];
  }
  __syncthreads();
  for (int c = warp; c < p.C; c += ATT_THREADS / 32) {
    const float* w = p.fc_w + (size_t)c * H;
    float s = 0.f;
    for (int h = lane; h < H; h += 32) s += __ldg(w + h) * o[h];
    s = warp_sum(s);
    if (lane == 0) {
      s += p.fc_b[c];
      lg[c] = s;
      if (p.logits) p.logits[(size_t)b * p.logit_stride + c] = s;
    }
  }
  __syncthreads();
  if (warp == 0 && (p.logp || p.argmax || p.history || p.end_count)) {
    float mx = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < p.C; c += 32)
      if (lg[c] > mx) { mx = lg[c]; am = c; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, off);
      const int oa = __shfl_xor_sync(0xffffffffu, am, off);
      if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
    if (p.logp) {
      float se = 0.f;
      for (int c = lane; c < p.C; c += 32) se += expf(lg[c] - mx);
      se = warp_sum(se);
      const float lse = mx + logf(se);
      for (int c = lane; c < p.C; c += 32) p.logp[(size_t)b * p.C + c] = lg[c] - lse;
    }
    if (lane == 0) {
      if (p.argmax) p.argmax[b] = am;
      if (p.history) p.history[(size_t)b * p.hist_stride + p.hist_col] = am;
      if (p.end_count && am == p.end_tok) atomicAdd(p.end_count, 1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention + output projection backward (one CTA per utterance)
// ------------------------------------------------------------------------------------------------
struct AttnBwdParams {
  const float* eh;         // (B, T, H)
  const float* hx;         // (B, H) decoder state of this step
  const float* hx_prev;    // (B, H)
  const float* ax_prev;    // (B, T) or null (first step)
  const float* ax;         // (B, T) this step's alignment (saved by forward)
  const float* sx;         // (B, H) this step's context (saved by forward)
  const float* conv_w; const float* conv_b; const float* lin_w; float lin_b;
  const float* fc_w;       // (C, H)
  const float* dlogits;    // dlogits[b*dl_stride + c]: gradient of the loss w.r.t. this step's logits
  long long dl_stride;
  const float* d_ix_next;  // (B, H) gradient w.r.t. the NEXT step's ix (= d sx through ix = emb + sx); null at the last step
  const float* d_ax_next;  // (B, T) gradient w.r.t. ax from the next step's conv; null at the last step
  const float* d_hx_next;  // (B, H) gradient w.r.t. hx from the next step's cell; null at the last step
  const float* gates;      // (B, 4, H) r, z, n, hn saved by the cell
  float* d_eh;             // (B, T, H) += 
  float* d_ax_prev;        // (B, T) out (gradient w.r.t. the previous alignment), unused at the first step
  float* d_gi;             // (B, 3H) out: gate pre-activation gradients of the cell (input side)
  float* d_gh;             // (B, 3H) out: (hidden side: the n entry carries r)
  float* d_hx_direct;      // (B, H) out: z * d hx (direct path to hx_prev)
  float* o_save;           // (B, H) out: hx + sx (operand of the time-batched d fc_w)
  // per-utterance parameter gradients, accumulated over steps, reduced over b by the caller
  float* g_conv_w;         // (B, H, Kc) +=
  float* g_conv_b;         // (B, H) +=
  float* g_lin_w;          // (B, H) +=
  float* g_lin_b;          // (B) +=
  int B, T, H, Kc, C, log_t;
};

__global__ void __launch_bounds__(ATT_THREADS) s2s_attn_bwd_kernel(const AttnBwdParams p) {
  extern __shared__ float bw_smem[];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, H = p.H, Kc = p.Kc, pad = (Kc - 1) / 2;
  constexpr int NW = ATT_THREADS / 32;
  float* axp = bw_smem;                        // [T + Kc - 1] zero-padded previous alignment
  float* daxp = axp + T + Kc - 1;              // [T + Kc - 1] gradient w.r.t. the padded alignment
  float* cw = daxp + T + Kc - 1;               // [H * Kc]
  float* dc = cw + H * Kc;                     // [H] hx + conv bias
  float* lw = dc + H;                          // [H]
  float* dsx = lw + H;                         // [H] total gradient w.r.t. sx
  float* dsc = dsx + H;                        // [T] da_t, then d score_t
  float* red = dsc + T;                        // [NW]
  float* dhx_att = red + NW;                   // [H] sum_t d pre[t, h]
  float* gcw = dhx_att + H;                    // [H * Kc] this step's conv weight gradient
  const bool has_prev = p.ax_prev != nullptr;
  const float* eh = p.eh + (size_t)b * T * H;
  const float* axv = p.ax + (size_t)b * T;
  const float tscale = p.log_t ? logf((float)T) : 1.0f;

  for (int k = tid; k < T + Kc - 1; k += ATT_THREADS) {
    const int t = k - pad;
    axp[k] = (has_prev && t >= 0 && t < T) ? p.ax_prev[(size_t)b * T + t] : 0.f;
    daxp[k] = 0.f;
  }
  if (has_prev)
    for (int k = tid; k < H * Kc; k += ATT_THREADS) { cw[k] = p.conv_w[k]; gcw[k] = 0.f; }
  // d o = W_fc^T dlogits ; o = hx + sx  ->  d sx = d o + d ix_next ; d hx gets d o as well
  for (int h = tid; h < H; h += ATT_THREADS) {
    float s = 0.f;
    for (int c = 0; c < p.C; ++c)
      s += p.dlogits[(size_t)b * p.dl_stride + c] * __ldg(p.fc_w + (size_t)c * H + h);
    dc[h] = p.hx[(size_t)b * H + h] + (has_prev ? p.conv_b[h] : 0.f);
    lw[h] = p.lin_w[h];
    dhx_att[h] = s;                                   // (d o; the attention part is added below)
    dsx[h] = s + (p.d_ix_next ? p.d_ix_next[(size_t)b * H + h] : 0.f);
    p.o_save[(size_t)b * H + h] = p.hx[(size_t)b * H + h] + p.sx[(size_t)b * H + h];
  }
  __syncthreads();
  // ---- pass 1: da_t = eh_t . d sx + d ax_next[t];  S = sum_t a_t da_t ----
  float spart = 0.f;
  for (int t = warp; t < T; t += NW) {
    float part = 0.f;
    for (int h = lane; h < H; h += 32) part += __ldg(eh + (size_t)t * H + h) * dsx[h];
    part = warp_sum(part);
    const float da = part + (p.d_ax_next ? p.d_ax_next[(size_t)b * T + t] : 0.f);
    if (lane == 0) dsc[t] = da;
    spart += axv[t] * da;
  }
  if (lane == 0) red[warp] = spart;
  __syncthreads();
  float S = 0.f;
  for (int w = 0; w < NW; ++w) S += red[w];
  __syncthreads();
  // d score_t = tscale * a_t (da_t - S)
  for (int t = tid; t < T; t += ATT_THREADS) dsc[t] = tscale * axv[t] * (dsc[t] - S);
  __syncthreads();
  // ---- pass 2a (warp per frame, lanes over h): through relu / linear / conv input; d eh ----
  float glw[ATT_MAXR], gdh[ATT_MAXR];
#pragma unroll
  for (int r = 0; r < ATT_MAXR; ++r) { glw[r] = 0.f; gdh[r] = 0.f; }
  float glb = 0.f;
  float* deh = p.d_eh + (size_t)b * T * H;
  for (int t = warp; t < T; t += NW) {
    const float ds = dsc[t];
    const float at = axv[t];
    if (lane == 0) glb += ds;
    float dk[16];                      // this lane's share of d axp[t + k]  (Kc <= 16)
#pragma unroll
    for (int k = 0; k < 16; ++k) dk[k] = 0.f;
#pragma unroll
    for (int r = 0; r < ATT_MAXR; ++r) {
      const int h = lane + 32 * r;
      if (h < H) {
        const float e = __ldg(eh + (size_t)t * H + h);
        float v = e + dc[h];
        const float* c = cw + h * Kc;
        if (has_prev) {
          float s = 0.f;
          for (int k = 0; k < Kc; ++k) s += c[k] * axp[t + k];
          v += s;
        }
        const float dpre = v > 0.f ? ds * lw[h] : 0.f;
        glw[r] += ds * fmaxf(v, 0.f);
        gdh[r] += dpre;
        deh[(size_t)t * H + h] += at * dsx[h] + dpre;
        if (has_prev) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k < Kc) dk[k] += dpre * c[k];
        }
      }
    }
    if (has_prev) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (k < Kc) {
          const float sk = warp_sum(dk[k]);
          if (lane == 0) atomicAdd(&daxp[t + k], sk);   // neighbouring frames overlap in t + k
        }
      }
    }
  }
  // ---- pass 2b (thread per h, all frames): conv weight gradient d cw[h, k] in registers ----
  if (has_prev) {
    for (int h = tid; h < H; h += ATT_THREADS) {
      const float* c = cw + h * Kc;
      float g[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) g[k] = 0.f;
      const float dch = dc[h], lwh = lw[h];
      for (int t = 0; t < T; ++t) {
        float v = __ldg(eh + (size_t)t * H + h) + dch;
        for (int k = 0; k < Kc; ++k) v += c[k] * axp[t + k];
        if (v > 0.f) {
          const float dpre = dsc[t] * lwh;
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (k < Kc) g[k] += dpre * axp[t + k];
        }
      }
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < Kc) gcw[h * Kc + k] = g[k];
    }
  }
  // reduce the per-warp partials of d lin_w, d hx (attention part), d lin_b over the 8 warps
  __syncthreads();
  float* tmp = cw;                 // [2][H] scratch (cw is no longer needed)
  for (int h = tid; h < 2 * H; h += ATT_THREADS) tmp[h] = 0.f;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ATT_MAXR; ++r) {
    const int h = lane + 32 * r;
    if (h < H) {
      atomicAdd(&tmp[h], glw[r]);
      atomicAdd(&tmp[H + h], gdh[r]);
    }
  }
  if (lane == 0) red[warp] = glb;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += red[w];
    p.g_lin_b[b] += s;
  }
  // ---- outputs: parameter gradients, d ax_prev, gate gradients of the cell ----
  const float* gt = p.gates + (size_t)b * 4 * H;
  for (int h = tid; h < H; h += ATT_THREADS) {
    p.g_lin_w[(size_t)b * H + h] += tmp[h];
    if (has_prev) p.g_conv_b[(size_t)b * H + h] += tmp[H + h];
    // total gradient w.r.t. hx_u: output projection + attention query + next step's cell
    const float dh = dhx_att[h] + tmp[H + h] + (p.d_hx_next ? p.d_hx_next[(size_t)b * H + h] : 0.f);
    const float r = gt[h], z = gt[H + h], n = gt[2 * H + h], hn = gt[3 * H + h];
    const float hp = p.hx_prev[(size_t)b * H + h];
    const float dn = dh * (1.f - z) * (1.f - n * n);
    const float dz = dh * (hp - n) * z * (1.f - z);
    const float dr = dn * hn * r * (1.f - r);
    float* gi = p.d_gi + (size_t)b * 3 * H;
    float* gh = p.d_gh + (size_t)b * 3 * H;
    gi[h] = dr; gi[H + h] = dz; gi[2 * H + h] = dn;
    gh[h] = dr; gh[H + h] = dz; gh[2 * H + h] = dn * r;
    p.d_hx_direct[(size_t)b * H + h] = dh * z;
  }
  if (has_prev) {
    for (int k = tid; k < H * Kc; k += ATT_THREADS) p.g_conv_w[(size_t)b * H * Kc + k] += gcw[k];
    for (int t = tid; t < T; t += ATT_THREADS) p.d_ax_prev[(size_t)b * T + t] = daxp[t + pad];
  }
}

// ------------------------------------------------------------------------------------------------
// cell backward: d ix = d gi W_ih ; d hx_prev = d gh W_hh + d_hx_direct   (lanes over the batch)
// ------------------------------------------------------------------------------------------------
struct CellBwdParams {
  const float* d_gi;        // (B, 3H)
  const float* d_gh;        // (B, 3H)
  const float* d_hx_direct; // (B, H)
  const float* w_ih;        // [3H][H]
  const float* w_hh;        // [3H][H]
  float* d_ix;              // (B, H) out
  float* d_hx_prev;         // (B, H) out
  int B, H;
};

__global__ void __launch_bounds__(32 * S2S_UPC) s2s_cell_bwd_kernel(const CellBwdParams p) {
  extern __shared__ float cb_smem[];
  const int H = p.H, B = p.B, N3 = 3 * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int k = blockIdx.x * S2S_UPC + warp;            // output column (input / hidden unit)
  const int ld = S2S_KC + 4;
  float* gis = cb_smem;                                 // [32][ld] d gi chunk
  float* ghs = cb_smem + 32 * ld;                       // [32][ld] d gh chunk
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int b = b0 + lane;
    float ai = 0.f, ah = 0.f;
    for (int n0 = 0; n0 < N3; n0 += S2S_KC) {
      const int nc = min(S2S_KC, N3 - n0);
      __syncthreads();
      for (int e = tid; e < 32 * (nc / 4); e += 32 * S2S_UPC) {
        const int r = e / (nc / 4), c4 = e % (nc / 4);
        const int rb = b0 + r;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
        if (rb < B) {
          a = __ldg(reinterpret_cast<const float4*>(p.d_gi + (long long)rb * N3 + n0) + c4);
          c = __ldg(reinterpret_cast<const float4*>(p.d_gh + (long long)rb * N3 + n0) + c4);
        }
        reinterpret_cast<float4*>(gis + r * ld)[c4] = a;
        reinterpret_cast<float4*>(ghs + r * ld)[c4] = c;
      }
      __syncthreads();
      if (k < H) {
        const float4* gr = reinterpret_cast<const float4*>(gis + lane * ld);
        const float4* hr = reinterpret_cast<const float4*>(ghs + lane * ld);
        // column k of W: consecutive n are H floats apart; the 8 warps of the CTA read 8
        // adjacent columns of the same 32-byte sector (broadcast over the lanes)
        const float* wi = p.w_ih + (long long)n0 * H + k;
        const float* wh = p.w_hh + (long long)n0 * H + k;
#pragma unroll 2
        for (int n4 = 0; n4 < nc / 4; ++n4) {
          const float4 a = gr[n4], c = hr[n4];
          const long long o = (long long)n4 * 4 * H;
          ai += a.x * __ldg(wi + o) + a.y * __ldg(wi + o + H) + a.z * __ldg(wi + o + 2 * H) +
                a.w * __ldg(wi + o + 3 * H);
          ah += c.x * __ldg(wh + o) + c.y * __ldg(wh + o + H) + c.z * __ldg(wh + o + 2 * H) +
                c.w * __ldg(wh + o + 3 * H);
        }
      }
    }
    if (k < H && b < B) {
      p.d_ix[(long long)b * H + k] = ai;
      p.d_hx_prev[(long long)b * H + k] = ah + p.d_hx_direct[(long long)b * H + k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// beam search bookkeeping on the device (Seq2Seq.beam_search, seq2seq.py:180-227; one utterance,
// the beam entries are the rows of the step kernels).  Per step, ONE CTA:
//   candidates (i, c): score_i + logp[i][c] for every live row i (float64, like the reference's
//   Python floats), "sorted" by (score desc, i*C + c asc) = the reference's stable descending sort
//   over its (beam-outer, vocabulary-inner) candidate list (:200-204); the first K of them that
//   end in end_tok join `complete` (:207-209); the first K non-ended ones are the next beam
//   (:211-212, with the py3 list() fix); stop when the beam is empty (:214) or K completed
//   hypotheses beat the best live one (:217-221).
// Hypotheses are nodes (parent node, token); the best complete (else best live) hypothesis is
// back-tracked on the device when the search stops.
// ------------------------------------------------------------------------------------------------
static constexpr int BM_MAXK = 32;
struct BeamState {
  double score[BM_MAXK];     // live beam scores
  int node[BM_MAXK];         // node id of each live entry
  int token[BM_MAXK];        // last token of each live entry (input of the next step)
  int nlive;
  int ncomplete;
  double best_c_score;       // best complete hypothesis (first inserted among equals)
  int best_c_node;
  int have_complete;
  int done;
  int nodes_used;
  int out_len;
};

struct BeamParams {
  const float* logp;         // (K, C) log-softmax of this step's logits
  BeamState* st;
  double* c_scores;          // [max complete] scores of complete hypotheses (for the stop rule)
  int* nodes;                // [node_cap][2] parent, token
  int* parent_row;           // (K) out: row of the previous beam each new entry continues
  int* tok_next;             // (K) out: token fed to the next step
  int* out_tokens;           // [max_len + 2] final hypothesis (written when the search stops)
  int K, C, end_tok, step, max_len, node_cap, c_cap;
};

__global__ void __launch_bounds__(256) s2s_beam_select_kernel(const BeamParams p) {
  extern __shared__ unsigned char bm_smem[];
  double* sc = reinterpret_cast<double*>(bm_smem);       // [K * C]
  __shared__ double rs[8];
  __shared__ int ri[8];
  __shared__ int sel_idx[2 * BM_MAXK];
  __shared__ double sel_sc[2 * BM_MAXK];
  const int tid = threadIdx.x;
  BeamState* st = p.st;
  if (st->done) return;
  const int K = p.K, C = p.C, nl = st->nlive;
  for (int i = tid; i < K * C; i += 256) {
    const int r = i / C, c = i - r * C;
    sc[i] = r < nl ? st->score[r] + (double)p.logp[(size_t)r * C + c] : nan("");
  }
  __syncthreads();
  const int want = min(2 * K, nl * C);
  for (int q = 0; q < want; ++q) {
    double bs = 0.0;
    int bi = -1;
    for (int i = tid; i < K * C; i += 256) {
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
    if ((tid & 31) == 0) { rs[tid >> 5] = bs; ri[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (ri[w] >= 0 && (bi < 0 || rs[w] > bs || (rs[w] == bs && ri[w] < bi))) { bs = rs[w]; bi = ri[w]; }
      sel_idx[q] = bi;
      sel_sc[q] = bs;
      if (bi >= 0) sc[bi] = nan("");
    }
    __syncthreads();
  }
  if (tid != 0) return;
  // ---- the reference's bookkeeping, sequential (a handful of entries) ----
  int old_node[BM_MAXK];
  for (int r = 0; r < nl; ++r) old_node[r] = st->node[r];
  int nb = 0;
  for (int q = 0; q < want; ++q) {
    const int i = sel_idx[q];
    if (i < 0) break;
    const int r = i / C, c = i - r * C;
    const bool ended = (c == p.end_tok);
    if (ended && q >= K) continue;                 // only the first K candidates may complete
    if (!ended && nb >= K) continue;
    int id = st->nodes_used;
    if (id >= p.node_cap) { st->done = 1; break; }
    st->nodes_used = id + 1;
    p.nodes[2 * id] = old_node[r];
    p.nodes[2 * id + 1] = c;
    if (ended) {
      if (st->ncomplete < p.c_cap) p.c_scores[st->ncomplete] = sel_sc[q];
      st->ncomplete += 1;
      if (!st->have_complete || sel_sc[q] > st->best_c_score) {
        st->have_complete = 1;
        st->best_c_score = sel_sc[q];
        st->best_c_node = id;
      }
    } else {
      st->score[nb] = sel_sc[q];
      st->node[nb] = id;
      st->token[nb] = c;
      p.parent_row[nb] = r;
      p.tok_next[nb] = c;
      ++nb;
    }
  }
  st->nlive = nb;
  bool stop = (nb == 0) || (p.step + 1 >= p.max_len);
  if (!stop) {
    int better = 0;
    const int nc = min(st->ncomplete, p.c_cap);
    for (int j = 0; j < nc; ++j) better += (p.c_scores[j] > st->score[0]) ? 1 : 0;
    stop = better >= K;
  }
  if (stop) {
    st->done = 1;
    // best complete hypothesis, else the best live one (seq2seq.py:223-227)
    int n = st->have_complete ? st->best_c_node : (nb > 0 ? st->node[0] : -1);
    int len = 0;
    for (int q = n; q >= 0; q = p.nodes[2 * q]) ++len;
    int k = len;
    for (int q = n; q >= 0; q = p.nodes[2 * q]) p.out_tokens[--k] = p.nodes[2 * q + 1];
    st->out_len = len;
  }
}

// rows of the next beam continue rows parent_row[] of the previous one: gather hx / sx / ax
__global__ void __launch_bounds__(256)
s2s_beam_gather_kernel(const float* __restrict__ hx_in, const float* __restrict__ sx_in,
                       const float* __restrict__ ax_in, float* __restrict__ hx_out,
                       float* __restrict__ sx_out, float* __restrict__ ax_out,
                       const int* __restrict__ parent_row, const BeamState* st, int H, int T) {
  if (st->done) return;
  const int r = blockIdx.x;
  if (r >= st->nlive) return;
  const int src = parent_row[r];
  for (int h = threadIdx.x; h < H; h += 256) {
    hx_out[(size_t)r * H + h] = hx_in[(size_t)src * H + h];
    sx_out[(size_t)r * H + h] = sx_in[(size_t)src * H + h];
  }
  for (int t = threadIdx.x; t < T; t += 256) ax_out[(size_t)r * T + t] = ax_in[(size_t)src * T + t];
}

// greedy decode: stop when EVERY row emitted end_tok at this step (seq2seq.py:162-178, :155-156)
__global__ void s2s_check_done_kernel(const int* end_count, int B, int* done, int* nsteps, int step1) {
  if (*done) return;
  *nsteps = step1;
  if (*end_count == B) *done = 1;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_s2s_cell_fwd(const float* emb, const int* tok, int tok_stride, const float* sx_prev,
                               const float* hx_prev, const float* w_ih, const float* w_hh,
                               const float* b_ih, const float* b_hh, float* hx, float* ix_save,
                               float* gates_save, const int* done, int B, int H, void* stream_) {
  if (!emb || !tok || !hx_prev || !w_ih || !w_hh || !b_ih || !b_hh || !hx || B <= 0 || H <= 0)
    return SB_ERR_INVALID;
  if (H % 4 != 0) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CellFwdParams p;
  p.emb = emb; p.tok = tok; p.tok_stride = tok_stride; p.sx_prev = sx_prev; p.hx_prev = hx_prev;
  p.w_ih = w_ih; p.w_hh = w_hh; p.b_ih = b_ih; p.b_hh = b_hh; p.hx = hx; p.ix_save = ix_save;
  p.gates_save = gates_save; p.done = done; p.B = B; p.H = H;
  const size_t smem = (size_t)2 * 32 * (S2S_KC + 4) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(s2s_cell_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return SB_ERR_CUDA;
    attr = true;
  }
  s2s_cell_fwd_kernel<<<(H + S2S_UPC - 1) / S2S_UPC, 32 * S2S_UPC, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_s2s_attn_fwd(const float* eh, int eh_bcast, const float* hx, const float* ax_prev,
                               const float* conv_w, const float* conv_b, const float* lin_w,
                               float lin_b, int log_t, int B, int T, int H, int Kc, float* sx,
                               float* ax, const float* fc_w, const float* fc_b, int C,
                               float* logits, long long logit_stride, float* logp, int* argmax,
                               int* history, int hist_stride, int hist_col, int* end_count,
                               int end_tok, const int* done, void* stream_) {
  if (!eh || !hx || !conv_w || !conv_b || !lin_w || !sx || !ax) return SB_ERR_INVALID;
  if (B <= 0 || T <= 0 || H <= 0 || Kc <= 0 || (Kc & 1) == 0) return SB_ERR_INVALID;
  if (H > 32 * ATT_MAXR) return SB_ERR_UNSUPPORTED;
  if (fc_w && (!fc_b || C <= 0 || C > (ATT_THREADS / 32 - 1) * H)) return SB_ERR_INVALID;
  const size_t smem = sizeof(float) * ((size_t)T + Kc - 1 + (size_t)H * Kc + 2 * H + T + 16 +
                                       (size_t)(ATT_THREADS / 32) * H);
  if (smem > 220 * 1024) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute(s2s_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem) != cudaSuccess)
    return SB_ERR_CUDA;
  AttnFwdParams p;
  p.eh = eh; p.eh_bcast = eh_bcast; p.hx = hx; p.ax_prev = ax_prev; p.conv_w = conv_w;
  p.conv_b = conv_b; p.lin_w = lin_w; p.lin_b = lin_b; p.sx = sx; p.ax = ax; p.fc_w = fc_w;
  p.fc_b = fc_b; p.logits = logits; p.logit_stride = logit_stride; p.logp = logp;
  p.argmax = argmax; p.history = history; p.hist_stride = hist_stride; p.hist_col = hist_col;
  p.end_count = end_count; p.end_tok = end_tok; p.done = done;
  p.B = B; p.T = T; p.H = H; p.Kc = Kc; p.C = C; p.log_t = log_t;
  s2s_attn_fwd_kernel<<<B, ATT_THREADS, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

// the standalone attention step (no output projection): NNAttention.forward on the decode path
extern "C" int sb_attn_step(const float* eh, const float* dhx, const float* ax_prev,
                            const float* conv_w, const float* conv_b, const float* lin_w,
                            float lin_b, int log_t, int B, int T, int H, int Kc, float* sx,
                            float* ax, void* stream_) {
  return sb_s2s_attn_fwd(eh, 0, dhx, ax_prev, conv_w, conv_b, lin_w, lin_b, log_t, B, T, H, Kc, sx,
                         ax, nullptr, nullptr, 0, nullptr, 0, nullptr, nullptr, nullptr, 0, 0,
                         nullptr, 0, nullptr, stream_);
}

extern "C" int sb_s2s_attn_bwd(const float* eh, const float* hx, const float* hx_prev,
                               const float* ax_prev, const float* ax, const float* sx,
                               const float* conv_w, const float* conv_b, const float* lin_w,
                               float lin_b, const float* fc_w, const float* dlogits,
                               long long dl_stride, const float* d_ix_next, const float* d_ax_next,
                               const float* d_hx_next, const float* gates, float* d_eh,
                               float* d_ax_prev, float* d_gi, float* d_gh, float* d_hx_direct,
                               float* o_save, float* g_conv_w, float* g_conv_b, float* g_lin_w,
                               float* g_lin_b, int log_t, int B, int T, int H, int Kc, int C,
                               void* stream_) {
  if (!eh || !hx || !hx_prev || !ax || !sx || !conv_w || !conv_b || !lin_w || !fc_w || !dlogits ||
      !gates || !d_eh || !d_ax_prev || !d_gi || !d_gh || !d_hx_direct || !o_save || !g_conv_w ||
      !g_conv_b || !g_lin_w || !g_lin_b)
    return SB_ERR_INVALID;
  if (B <= 0 || T <= 0 || H <= 0 || Kc <= 0 || (Kc & 1) == 0 || C <= 0) return SB_ERR_INVALID;
  if (H > 32 * ATT_MAXR || Kc > 16 || Kc < 2) return SB_ERR_UNSUPPORTED;   // (tmp = cw needs Kc >= 2)
  const size_t smem = sizeof(float) * (2 * ((size_t)T + Kc - 1) + 2 * (size_t)H * Kc + 4 * H + T +
                                       ATT_THREADS / 32 + 16);
  if (smem > 220 * 1024) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute(s2s_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem) != cudaSuccess)
    return SB_ERR_CUDA;
  AttnBwdParams p;
  p.eh = eh; p.hx = hx; p.hx_prev = hx_prev; p.ax_prev = ax_prev; p.ax = ax; p.sx = sx;
  p.conv_w = conv_w; p.conv_b = conv_b; p.lin_w = lin_w; p.lin_b = lin_b; p.fc_w = fc_w;
  p.dlogits = dlogits; p.dl_stride = dl_stride; p.d_ix_next = d_ix_next; p.d_ax_next = d_ax_next;
  p.d_hx_next = d_hx_next; p.gates = gates; p.d_eh = d_eh; p.d_ax_prev = d_ax_prev; p.d_gi = d_gi;
  p.d_gh = d_gh; p.d_hx_direct = d_hx_direct; p.o_save = o_save; p.g_conv_w = g_conv_w;
  p.g_conv_b = g_conv_b; p.g_lin_w = g_lin_w; p.g_lin_b = g_lin_b;
  p.B = B; p.T = T; p.H = H; p.Kc = Kc; p.C = C; p.log_t = log_t;
  s2s_attn_bwd_kernel<<<B, ATT_THREADS, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_s2s_cell_bwd(const float* d_gi, const float* d_gh, const float* d_hx_direct,
                               const float* w_ih, const float* w_hh, float* d_ix, float* d_hx_prev,
                               int B, int H, void* stream_) {
  if (!d_gi || !d_gh || !d_hx_direct || !w_ih || !w_hh || !d_ix || !d_hx_prev || B <= 0 || H <= 0)
    return SB_ERR_INVALID;
  if ((3 * H) % 4 != 0) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CellBwdParams p;
  p.d_gi = d_gi; p.d_gh = d_gh; p.d_hx_direct = d_hx_direct; p.w_ih = w_ih; p.w_hh = w_hh;
  p.d_ix = d_ix; p.d_hx_prev = d_hx_prev; p.B = B; p.H = H;
  const size_t smem = (size_t)2 * 32 * (S2S_KC + 4) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(s2s_cell_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return SB_ERR_CUDA;
    attr = true;
  }
  s2s_cell_bwd_kernel<<<(H + S2S_UPC - 1) / S2S_UPC, 32 * S2S_UPC, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_s2s_check_done(const int* end_count, int B, int* done, int* nsteps, int step1,
                                 void* stream_) {
  if (!end_count || !done || !nsteps) return SB_ERR_INVALID;
  s2s_check_done_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(end_count, B, done,
                                                                               nsteps, step1);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_s2s_beam_state_size(size_t* bytes) {
  if (!bytes) return SB_ERR_INVALID;
  *bytes = sizeof(sb::BeamState);
  return SB_OK;
}

// state must be zero-filled except: score[0] = 0, node[0] = root node id 0 (nodes[0] = {-1, start
// token}), token[0] = start token, nlive = 1, nodes_used = 1 -- sb_s2s_beam_init does that.
extern "C" int sb_s2s_beam_init(void* state, int* nodes, int* tok_next, int start_tok,
                                void* stream_) {
  if (!state || !nodes || !tok_next) return SB_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  sb::BeamState h;
  memset(&h, 0, sizeof(h));
  h.score[0] = 0.0; h.node[0] = 0; h.token[0] = start_tok; h.nlive = 1; h.nodes_used = 1;
  const int root[2] = {-1, start_tok};
  if (cudaMemcpyAsync(state, &h, sizeof(h), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
      cudaMemcpyAsync(nodes, root, sizeof(root), cudaMemcpyHostToDevice, stream) != cudaSuccess ||
      cudaMemcpyAsync(tok_next, &start_tok, sizeof(int), cudaMemcpyHostToDevice, stream) !=
          cudaSuccess)
    return SB_ERR_CUDA;
  // (pageable host sources: the copies are staged before the call returns)
  return SB_OK;
}

extern "C" int sb_s2s_beam_select(const float* logp, void* state, double* c_scores, int* nodes,
                                  int* parent_row, int* tok_next, int* out_tokens, int K, int C,
                                  int end_tok, int step, int max_len, int node_cap, int c_cap,
                                  void* stream_) {
  if (!logp || !state || !c_scores || !nodes || !parent_row || !tok_next || !out_tokens)
    return SB_ERR_INVALID;
  if (K <= 0 || K > sb::BM_MAXK || C <= 0) return SB_ERR_UNSUPPORTED;
  const size_t smem = (size_t)K * C * sizeof(double);
  if (smem > 200 * 1024) return SB_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute(sb::s2s_beam_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem) != cudaSuccess)
    return SB_ERR_CUDA;
  sb::BeamParams p;
  p.logp = logp; p.st = reinterpret_cast<sb::BeamState*>(state); p.c_scores = c_scores;
  p.nodes = nodes; p.parent_row = parent_row; p.tok_next = tok_next; p.out_tokens = out_tokens;
  p.K = K; p.C = C; p.end_tok = end_tok; p.step = step; p.max_len = max_len;
  p.node_cap = node_cap; p.c_cap = c_cap;
  sb::s2s_beam_select_kernel<<<1, 256, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

extern "C" int sb_s2s_beam_gather(const float* hx_in, const float* sx_in, const float* ax_in,
                                  float* hx_out, float* sx_out, float* ax_out,
                                  const int* parent_row, const void* state, int K, int H, int T,
                                  void* stream_) {
  if (!hx_in || !sx_in || !ax_in || !hx_out || !sx_out || !ax_out || !parent_row || !state)
    return SB_ERR_INVALID;
  sb::s2s_beam_gather_kernel<<<K, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      hx_in, sx_in, ax_in, hx_out, sx_out, ax_out, parent_row,
      reinterpret_cast<const sb::BeamState*>(state), H, T);
  return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
