"""Drop-in for `transducer.decoders` of awni/transducer (imported at
speech/models/transducer_model.py:10, used at :100 as `td.decode_static(lp, beam_size, blank)[0]`).

The dependency is un-vendored (Makefile:10-12) so its exact algorithm and tie-breaks are
UNVERIFIABLE; this is the standard transducer beam search (Graves 2012, §3)
restricted to a STATIC lattice: because `lp[t, u, :]` was computed with teacher forcing, the
prediction-network state of a hypothesis is simply the number of labels it has emitted, so
hypotheses are (label prefix, log-probability) pairs advancing through (t, u).
Returns (labels, log_probability) so that `[0]` is the label list, as the call site expects.

Beam bookkeeping is host-side (a handful of scalars per frame); the lattice stays on the device
and each frame's (U x V) slice is read back once.
"""
import math

import torch


def decode_static(lp, beam_size, blank=0, max_symbols_per_frame=None):
    lp = torch.as_tensor(lp)
    T, U, V = lp.shape
    lat = lp.detach().float().cpu().numpy()
    beam = {(): 0.0}
    for t in range(T):
        done = {}
        frontier = dict(beam)
        # expand within the frame until every surviving hypothesis has emitted its blank
        for _ in range(U if max_symbols_per_frame is None else max_symbols_per_frame + 1):
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
                        h2 = hyp + (k,)
                        s2 = score + float(row[k])
                        nxt[h2] = _lse(nxt[h2], s2) if h2 in nxt else s2
            if not nxt:
                break
            frontier = dict(sorted(nxt.items(), key=lambda kv: -kv[1])[:beam_size])
        beam = dict(sorted(done.items(), key=lambda kv: -kv[1])[:beam_size])
    best = max(beam.items(), key=lambda kv: kv[1])
    return list(best[0]), best[1]


def _lse(a, b):
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))
