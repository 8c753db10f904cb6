"""CPU oracle for `transducer.decoders.decode_static` (TEST INFRASTRUCTURE - never imported by
speech_b200/).

The reference calls `td.decode_static(lp, beam_size, blank=self.blank)[0]` on the TEACHER-FORCED
lattice of one utterance (speech/models/transducer_model.py:92-101); `transducer` is the
un-vendored awni/transducer @ master (Makefile:10-12), so its exact algorithm and tie-breaks are
UNVERIFIABLE - PARITY UNPINNED by the reference.  This restates the standard
transducer beam search (Graves 2012, section 3) on a static lattice: `lp[t, u, :]` was computed
with teacher forcing, so the prediction-network state of a hypothesis is the number of labels it
has emitted and hypotheses are (label prefix, log-probability) pairs advancing through (t, u).

It is pinned by `best_by_enumeration`: with a beam wide enough to hold every hypothesis the search
is exact, so it must return the label sequence of maximal total probability, which is computed
there from the definition (forward algorithm of every candidate sequence).
"""
import itertools
import math

import numpy as np


def _lse(a, b):
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def decode_static(lp, beam_size, blank=0):
    """lp (T, U, V) log-probabilities -> (labels, log-probability of the best hypothesis)."""
    lat = np.asarray(lp, dtype=np.float32)
    T, U, V = lat.shape
    beam = {(): 0.0}
    for t in range(T):
        done = {}
        frontier = dict(beam)
        for _ in range(U):      # a frame can emit at most U-1 labels before its blank
            nxt = {}
            for hyp, score in frontier.items():
                u = len(hyp)
                if u >= U:
                    continue
                row = lat[t, u]
                b = score + float(row[blank])
                done[hyp] = _lse(done[hyp], b) if hyp in done else b
                if u + 1 < U:
                    for k in range(V):
                        if k == blank:
                            continue
                        nxt[hyp + (k,)] = score + float(row[k])
            if not nxt:
                break
            frontier = dict(sorted(nxt.items(), key=lambda kv: -kv[1])[:beam_size])
        beam = dict(sorted(done.items(), key=lambda kv: -kv[1])[:beam_size])
    best = max(beam.items(), key=lambda kv: kv[1])
    return list(best[0]), best[1]


def sequence_log_prob(lp, labels, blank):
    """log of the total probability of emitting exactly `labels` over the static lattice."""
    lat = np.asarray(lp, dtype=np.float64)
    T = lat.shape[0]
    U = len(labels)
    alpha = np.full((T, U + 1), -np.inf)
    for t in range(T):
        for u in range(U + 1):
            if t == 0 and u == 0:
                alpha[t, u] = 0.0
                continue
            a = alpha[t - 1, u] + lat[t - 1, u, blank] if t > 0 else -np.inf
            c = alpha[t, u - 1] + lat[t, u - 1, labels[u - 1]] if u > 0 else -np.inf
            alpha[t, u] = np.logaddexp(a, c)
    return alpha[T - 1, U] + lat[T - 1, U, blank]


def best_by_enumeration(lp, blank):
    """arg-max over EVERY label sequence the lattice admits (length < U): tiny lattices only."""
    lat = np.asarray(lp, dtype=np.float64)
    T, U, V = lat.shape
    symbols = [k for k in range(V) if k != blank]
    best, best_lp = None, -np.inf
    for n in range(U):
        for seq in itertools.product(symbols, repeat=n):
            s = sequence_log_prob(lat, list(seq), blank)
            if s > best_lp:
                best, best_lp = list(seq), s
    return best, best_lp
