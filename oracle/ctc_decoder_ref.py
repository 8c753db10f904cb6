"""CPU oracle for the CTC prefix beam search (TEST INFRASTRUCTURE - never imported by speech_b200/).

Restates speech/models/ctc_decoder.py:38-113 of the reference:
  - input is post-softmax probabilities (T, S); their log is taken in the INPUT dtype (:52);
  - every beam entry is a prefix with two log-masses: paths ending in blank / in a non-blank;
  - a blank keeps the prefix (:75-79); a symbol extends it, and a repeated symbol only extends
    from the blank-ending mass (:84-96) while also feeding the un-extended prefix (:100-103);
  - candidates live in an insertion-ordered dict touched in `for s: for prefix:` order (:65,:71);
    pruning is a STABLE descending sort on logsumexp(p_b, p_nb) (:107-110), so ties keep
    first-touch order.
Lattice arithmetic is float64 (the reference's Python floats under its pinned numpy 1.13,
SURVEY.md §8c).  Pinned against the reference's own function by tests/golden/make_golden.py.
"""
import math

import numpy as np

NEG_INF = -float("inf")


def _lse(*xs):
    if all(x == NEG_INF for x in xs):
        return NEG_INF
    m = max(xs)
    return m + math.log(sum(math.exp(x - m) for x in xs))


def prefix_beam_search(probs, beam_size=10, blank=0):
    probs = np.asarray(probs)
    T, S = probs.shape
    with np.errstate(divide="ignore"):
        logp = np.log(probs)
    beam = [((), (0.0, NEG_INF))]
    for t in range(T):
        cand = {}          # prefix -> [p_b, p_nb]; dict preserves first-touch order

        def slot(prefix):
            if prefix not in cand:
                cand[prefix] = [NEG_INF, NEG_INF]
            return cand[prefix]

        for s in range(S):
            p = float(logp[t, s])
            for prefix, (p_b, p_nb) in beam:
                if s == blank:
                    e = slot(prefix)
                    e[0] = _lse(e[0], p_b + p, p_nb + p)
                    continue
                last = prefix[-1] if prefix else None
                e = slot(prefix + (s,))
                if s != last:
                    e[1] = _lse(e[1], p_b + p, p_nb + p)
                else:
                    e[1] = _lse(e[1], p_b + p)
                    k = slot(prefix)
                    k[1] = _lse(k[1], p_nb + p)
        ranked = sorted(cand.items(), key=lambda kv: _lse(*kv[1]), reverse=True)
        beam = [(k, (v[0], v[1])) for k, v in ranked[:beam_size]]
    best = beam[0]
    return best[0], -_lse(*best[1])
