"""CPU oracle for the RNN-Transducer loss + gradient (TEST INFRASTRUCTURE - never imported by
speech_b200/).

Restates what `transducer.functions.transducer.TransducerLoss` computes at the reference's call
site (speech/models/transducer_model.py:46-52): (B, T, U+1, V+1) log-probabilities, flat labels,
blank = last class, costs summed by the caller; gradient w.r.t. the log-probabilities.  The
arithmetic lives in the un-vendored awni/transducer @ master (Makefile:10-12), so this follows the
published algorithm (Graves 2012, eqs. 16-20) in float64.

PARITY UNPINNED by the reference (it has no transducer test at all); pinned instead
against `brute_force_nll`, which enumerates every alignment of tiny lattices from the definition.
"""
import itertools

import numpy as np

NEG = -np.inf


def _lse(a, b):
    m = max(a, b)
    if m == NEG:
        return NEG
    return m + np.log(np.exp(a - m) + np.exp(b - m))


def rnnt_single(lp, labels, blank):
    """lp (T, U+1, V) log-probs (only the first len(labels)+1 rows of axis 1 are used)."""
    lp = np.asarray(lp, np.float64)
    T = lp.shape[0]
    U = len(labels)
    grad = np.zeros_like(lp)
    if T == 0:
        return (0.0 if U == 0 else np.inf), grad
    alpha = np.full((T, U + 1), NEG)
    beta = np.full((T, U + 1), NEG)
    for t in range(T):
        for u in range(U + 1):
            if t == 0 and u == 0:
                alpha[t, u] = 0.0
                continue
            a = alpha[t - 1, u] + lp[t - 1, u, blank] if t > 0 else NEG
            c = alpha[t, u - 1] + lp[t, u - 1, labels[u - 1]] if u > 0 else NEG
            alpha[t, u] = _lse(a, c)
    for t in range(T - 1, -1, -1):
        for u in range(U, -1, -1):
            if t == T - 1 and u == U:
                beta[t, u] = lp[t, u, blank]
                continue
            a = beta[t + 1, u] + lp[t, u, blank] if t < T - 1 else NEG
            c = beta[t, u + 1] + lp[t, u, labels[u]] if u < U else NEG
            beta[t, u] = _lse(a, c)
    logp = beta[0, 0]
    if logp == NEG:
        return np.inf, grad
    for t in range(T):
        for u in range(U + 1):
            if alpha[t, u] == NEG:
                continue
            nxt = beta[t + 1, u] if t < T - 1 else (0.0 if u == U else NEG)
            if nxt != NEG:
                grad[t, u, blank] = -np.exp(alpha[t, u] + lp[t, u, blank] + nxt - logp)
            if u < U and beta[t, u + 1] != NEG:
                k = labels[u]
                grad[t, u, k] = -np.exp(alpha[t, u] + lp[t, u, k] + beta[t, u + 1] - logp)
    return -logp, grad


def rnnt_loss_and_grad(lp, labels_flat, x_lens, y_lens, blank=None):
    lp = np.asarray(lp, np.float64)
    B = lp.shape[0]
    if blank is None:
        blank = lp.shape[3] - 1
    costs = np.zeros(B)
    grads = np.zeros_like(lp)
    off = 0
    for b in range(B):
        U = int(y_lens[b])
        T = int(x_lens[b])
        c, g = rnnt_single(lp[b, :T], [int(v) for v in labels_flat[off:off + U]], blank)
        costs[b] = c
        grads[b, :T] = g
        off += U
    return costs, grads


def brute_force_nll(lp, labels, blank):
    """-log of the sum over every alignment (T blanks interleaved with the U labels, the last
    symbol being a blank emitted at t = T-1).  Exponential: tiny lattices only."""
    lp = np.asarray(lp, np.float64)
    T, U = lp.shape[0], len(labels)
    total = 0.0
    # an alignment = positions (time steps) at which each label is emitted, non-decreasing
    for times in itertools.combinations_with_replacement(range(T), U):
        logp = 0.0
        u = 0
        for t in range(T):
            while u < U and times[u] == t:
                logp += lp[t, u, labels[u]]
                u += 1
            logp += lp[t, u, blank]
        total += np.exp(logp)
    return -np.log(total)
