"""CPU oracle for the CTC loss + gradient (TEST INFRASTRUCTURE - never imported by speech_b200/).

Restates what `functions.ctc.CTCLoss` computes at the reference's call site
(speech/models/ctc_model.py:34-40): activations (B, T, V) batch-first un-normalised, softmax
internal, blank = last class (ctc_model.py:18), flat labels, cost summed over the minibatch.
The arithmetic itself lives in the un-vendored dependency awni/warp-ctc @ master
(Makefile:4-7; not present under /root/reference) so this follows the published algorithm
(Graves et al. 2006, eqs. 6-16) in float64 log space.

PARITY PINNING: the reference holds no golden CTC value (tests/ctc_test.py:26 discards the loss);
this oracle is pinned instead against torch.nn.functional.ctc_loss (an independent
implementation) in tests/test_oracle.py, and against the compiled C restatement oracle/ctc_ref.c.
"""
import numpy as np

NEG_INF = -np.inf


def _logsumexp(*xs):
    m = max(xs)
    if m == NEG_INF:
        return NEG_INF
    return m + np.log(sum(np.exp(x - m) for x in xs))


def log_softmax(x):
    x = np.asarray(x, dtype=np.float64)
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def ctc_single(logits, labels, blank):
    """One utterance.  logits (T, V) raw; labels list[int].  Returns (nll, grad (T, V)) float64."""
    lp = log_softmax(logits)
    T, V = lp.shape
    L = len(labels)
    S = 2 * L + 1
    ext = [blank] * S
    for i, l in enumerate(labels):
        ext[2 * i + 1] = int(l)
    grad = np.exp(lp)  # softmax; occupancy is subtracted below
    if T == 0:
        return (0.0 if L == 0 else np.inf), grad

    alpha = np.full((T, S), NEG_INF)
    alpha[0, 0] = lp[0, ext[0]]
    if S > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        for s in range(S):
            terms = [alpha[t - 1, s]]
            if s >= 1:
                terms.append(alpha[t - 1, s - 1])
            if s >= 2 and (s & 1) and ext[s] != ext[s - 2]:
                terms.append(alpha[t - 1, s - 2])
            alpha[t, s] = _logsumexp(*terms) + lp[t, ext[s]]

    beta = np.full((T, S), NEG_INF)
    beta[T - 1, S - 1] = lp[T - 1, ext[S - 1]]
    if S > 1:
        beta[T - 1, S - 2] = lp[T - 1, ext[S - 2]]
    for t in range(T - 2, -1, -1):
        for s in range(S):
            terms = [beta[t + 1, s]]
            if s + 1 < S:
                terms.append(beta[t + 1, s + 1])
            if s + 2 < S and (s & 1) and ext[s] != ext[s + 2]:
                terms.append(beta[t + 1, s + 2])
            beta[t, s] = _logsumexp(*terms) + lp[t, ext[s]]

    tail = [alpha[T - 1, S - 1]]
    if S > 1:
        tail.append(alpha[T - 1, S - 2])
    logp = _logsumexp(*tail)
    if logp == NEG_INF:
        return np.inf, np.zeros_like(grad)
    for t in range(T):
        occ = np.zeros(V)
        for s in range(S):
            v = alpha[t, s] + beta[t, s]
            if v != NEG_INF:
                occ[ext[s]] += np.exp(v - lp[t, ext[s]] - logp)
        grad[t] -= occ
    return -logp, grad


def ctc_loss_and_grad(acts, labels_flat, act_lens, label_lens, blank=None):
    """Minibatch.  acts (B,T,V).  Returns (costs (B,), grads (B,T,V)); rows past act_lens are 0."""
    acts = np.asarray(acts, dtype=np.float64)
    B, T, V = acts.shape
    if blank is None:
        blank = V - 1
    costs = np.zeros(B)
    grads = np.zeros_like(acts)
    off = 0
    for b in range(B):
        L = int(label_lens[b])
        Tb = int(act_lens[b])
        c, g = ctc_single(acts[b, :Tb], list(labels_flat[off:off + L]), blank)
        costs[b] = c
        grads[b, :Tb] = g
        off += L
    return costs, grads
