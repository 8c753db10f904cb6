/* oracle/ctc_ref.c - C restatement of the CTC loss + gradient (TEST INFRASTRUCTURE ONLY).
 *
 * Same contract as functions.ctc.CTCLoss at the reference's call site
 * (speech/models/ctc_model.py:34-40): activations (B,T,V) batch-first raw logits, softmax
 * internal, flat labels, per-utterance costs; arithmetic after Graves et al. 2006 (the
 * un-vendored awni/warp-ctc, Makefile:4-7, is not available).  float64 log-space, OpenMP over
 * the minibatch (like warp-ctc's CPU path).  Used to cross-check oracle/ctc_ref.py at sizes the
 * numpy loops cannot finish, and as a timed CPU leg of the CTC micro-benchmark.
 * Build: oracle/build.py (gcc -O2 -fopenmp -shared -fPIC) -> oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static double lse2(double a, double b) {
  double m = a > b ? a : b;
  if (m == -INFINITY) return -INFINITY;
  return m + log(exp(a - m) + exp(b - m));
}

static void one_utt(const float* acts, int T, int V, const int* lab, int L, int blank,
                    double* cost, float* grad) {
  int S = 2 * L + 1, t, s, k;
  if (T <= 0) { *cost = (L == 0) ? 0.0 : INFINITY; return; }
  double* lp = (double*)malloc(sizeof(double) * (size_t)T * V);
  double* al = (double*)malloc(sizeof(double) * (size_t)T * S);
  double* be = (double*)malloc(sizeof(double) * (size_t)T * S);
  int* ext = (int*)malloc(sizeof(int) * S);
  for (s = 0; s < S; ++s) ext[s] = (s & 1) ? lab[s / 2] : blank;
  for (t = 0; t < T; ++t) {
    double m = -INFINITY, z = 0.0;
    for (k = 0; k < V; ++k) if (acts[t * V + k] > m) m = acts[t * V + k];
    for (k = 0; k < V; ++k) z += exp(acts[t * V + k] - m);
    z = m + log(z);
    for (k = 0; k < V; ++k) lp[t * V + k] = acts[t * V + k] - z;
  }
  for (s = 0; s < T * S; ++s) { al[s] = -INFINITY; be[s] = -INFINITY; }
  al[0] = lp[ext[0]];
  if (S > 1) al[1] = lp[ext[1]];
  for (t = 1; t < T; ++t)
    for (s = 0; s < S; ++s) {
      double a = al[(t - 1) * S + s];
      if (s >= 1) a = lse2(a, al[(t - 1) * S + s - 1]);
      if (s >= 2 && (s & 1) && ext[s] != ext[s - 2]) a = lse2(a, al[(t - 1) * S + s - 2]);
      al[t * S + s] = a + lp[t * V + ext[s]];
    }
  be[(T - 1) * S + S - 1] = lp[(T - 1) * V + ext[S - 1]];
  if (S > 1) be[(T - 1) * S + S - 2] = lp[(T - 1) * V + ext[S - 2]];
  for (t = T - 2; t >= 0; --t)
    for (s = 0; s < S; ++s) {
      double b = be[(t + 1) * S + s];
      if (s + 1 < S) b = lse2(b, be[(t + 1) * S + s + 1]);
      if (s + 2 < S && (s & 1) && ext[s] != ext[s + 2]) b = lse2(b, be[(t + 1) * S + s + 2]);
      be[t * S + s] = b + lp[t * V + ext[s]];
    }
  double logp = al[(T - 1) * S + S - 1];
  if (S > 1) logp = lse2(logp, al[(T - 1) * S + S - 2]);
  *cost = -logp;
  if (grad) {
    if (logp == -INFINITY) {
      memset(grad, 0, sizeof(float) * (size_t)T * V);
    } else {
      double* occ = (double*)malloc(sizeof(double) * V);
      for (t = 0; t < T; ++t) {
        for (k = 0; k < V; ++k) occ[k] = 0.0;
        for (s = 0; s < S; ++s) {
          double v = al[t * S + s] + be[t * S + s];
          if (v != -INFINITY) occ[ext[s]] += exp(v - lp[t * V + ext[s]] - logp);
        }
        for (k = 0; k < V; ++k) grad[t * V + k] = (float)(exp(lp[t * V + k]) - occ[k]);
      }
      free(occ);
    }
  }
  free(lp); free(al); free(be); free(ext);
}

/* acts (B,T,V) f32; grads (B,T,V) f32 or NULL; labels flat; returns 0 */
int oracle_ctc(const float* acts, float* grads, const int* labels, const int* label_lens,
               const int* act_lens, int B, int T, int V, int blank, double* costs) {
  int* off = (int*)malloc(sizeof(int) * (B + 1));
  int b;
  off[0] = 0;
  for (b = 0; b < B; ++b) off[b + 1] = off[b] + label_lens[b];
  if (grads) memset(grads, 0, sizeof(float) * (size_t)B * T * V);
#pragma omp parallel for schedule(dynamic)
  for (b = 0; b < B; ++b) {
    int Tb = act_lens[b] < T ? act_lens[b] : T;
    one_utt(acts + (size_t)b * T * V, Tb, V, labels + off[b], label_lens[b], blank, &costs[b],
            grads ? grads + (size_t)b * T * V : 0);
  }
  free(off);
  return 0;
}
