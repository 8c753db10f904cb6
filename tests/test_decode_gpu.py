"""GPU parity: sb_ctc_prefix_beam vs the reference's golden hypotheses and the CPU oracle.
Bar: label sequences identical; scores within 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc_decoder_ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_decode_matches_reference_golden(cuda_lib):
    from speech_b200.models.ctc_decoder import decode
    g = np.load(os.path.join(GOLD, "ctc_decode.npz"))
    cases = [("demo", 0, (10, 8, 1)), ("f32", 11, (1, 4, 10)), ("peaky", 4, (1, 3, 8)),
             ("tie", 0, (1, 3))]
    for name, blank, beams in cases:
        probs = g[name + "_probs"]
        for b in beams:
            lab, sc = decode(probs, beam_size=b, blank=blank)
            assert list(lab) == list(g["%s_b%d_labels" % (name, b)]), (name, b)
            ref = float(g["%s_b%d_score" % (name, b)])
            assert abs(sc - ref) < 1e-5 * max(1.0, abs(ref)), (name, b, sc, ref)


@pytest.mark.parametrize("T,S,beam,blank,seed", [(247, 29, 1, 28, 0), (247, 29, 8, 28, 1),
                                                 (33, 5, 32, 0, 2), (100, 62, 10, 61, 3)])
def test_decode_batch_matches_oracle(cuda_lib, T, S, beam, blank, seed):
    from speech_b200.models.ctc_decoder import decode_batch
    rng = np.random.RandomState(seed)
    B = 6
    logits = rng.randn(B, T, S).astype(np.float32) * 1.5
    probs = torch.softmax(torch.from_numpy(logits), 2)
    lens = np.array([T, T - 3, T, max(1, T // 2), T, 1], np.int32)
    res = decode_batch(probs.cuda(), beam_size=beam, blank=blank, lens=lens, with_scores=True)
    pn = probs.numpy()
    for b in range(B):
        lab, sc = ctc_decoder_ref.prefix_beam_search(pn[b, :lens[b]], beam_size=beam, blank=blank)
        assert tuple(lab) == res[b][0], b
        assert abs(sc - res[b][1]) < 1e-5 * max(1.0, abs(sc))
