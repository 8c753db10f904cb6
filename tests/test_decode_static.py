"""decode_static (transducer beam search on a static lattice).

CPU: the oracle restatement (oracle/decode_static_ref.py) is pinned by exhaustive enumeration: with
a beam that holds every hypothesis the search is exact, so it must return the label sequence of
maximal total probability computed from the definition.
GPU: the kernel (csrc/tdecode.cu) must return the oracle's hypotheses label for label, and its
score to 1e-9, on random lattices (several beam widths, peaky lattices that force prefix merging,
ragged T / U inside a batch)."""
import numpy as np
import pytest
import torch

from oracle import decode_static_ref as ref


def _lattice(rng, T, U, V, peak=0.0):
    x = rng.randn(T, U, V)
    if peak:
        x[..., rng.randint(0, V)] += peak
    x = x - np.log(np.exp(x).sum(-1, keepdims=True))
    return x.astype(np.float32)


@pytest.mark.parametrize("T,U,V,seed", [(2, 2, 3, 0), (3, 3, 3, 1), (4, 3, 4, 2), (3, 4, 3, 3),
                                        (5, 3, 3, 4)])
def test_oracle_with_a_wide_beam_equals_enumeration(T, U, V, seed):
    rng = np.random.RandomState(seed)
    lat = _lattice(rng, T, U, V)
    blank = V - 1
    labels, score = ref.decode_static(lat, beam_size=10 ** 6, blank=blank)
    best, best_lp = ref.best_by_enumeration(lat, blank)
    assert labels == best
    assert abs(score - best_lp) < 1e-6
    # and the score of ANY returned hypothesis is its exact total log-probability
    l2, s2 = ref.decode_static(lat, beam_size=2, blank=blank)
    assert abs(s2 - ref.sequence_log_prob(lat, l2, blank)) < 1e-6 or s2 <= best_lp + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("beam", [1, 2, 4, 8])
@pytest.mark.parametrize("T,U,V,peak,seed", [(6, 4, 5, 0.0, 0), (20, 8, 11, 0.0, 1),
                                             (30, 12, 29, 3.0, 2), (48, 21, 11, 2.0, 3),
                                             (12, 3, 62, 0.0, 4)])
def test_kernel_matches_oracle(cuda_lib, beam, T, U, V, peak, seed):
    from speech_b200.transducer_decoders import decode_static
    rng = np.random.RandomState(seed)
    lat = _lattice(rng, T, U, V, peak)
    for blank in (V - 1, 0):
        want, wscore = ref.decode_static(lat, beam, blank)
        got, gscore = decode_static(lat, beam, blank)
        assert got == want, (beam, blank)
        assert abs(gscore - wscore) < 1e-9 * max(1.0, abs(wscore))


@pytest.mark.gpu
def test_batched_ragged_lattices(cuda_lib):
    from speech_b200.transducer_decoders import decode_static_batch
    rng = np.random.RandomState(7)
    B, T, U1, V = 5, 25, 9, 11
    lp = torch.from_numpy(np.stack([_lattice(rng, T, U1, V, 1.5) for _ in range(B)]))
    tlens = [25, 17, 25, 3, 1]
    ulens = [9, 9, 4, 2, 1]
    labels, scores = decode_static_batch(lp.cuda(), tlens, ulens, 4, V - 1)
    for b in range(B):
        want, ws = ref.decode_static(lp[b, :tlens[b], :ulens[b]].numpy(), 4, V - 1)
        assert labels[b] == want, b
        assert abs(scores[b] - ws) < 1e-9 * max(1.0, abs(ws))
        assert len(labels[b]) <= ulens[b] - 1
