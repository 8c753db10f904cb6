"""GPU parity of the featuriser kernel (csrc/specgram.cu through sb_log_specgram) against the
float64 oracle restatement of the reference's `log_specgram` (oracle/specgram_ref.py) and against
the golden output of the reference's own function on its own wav fixtures.

Tolerances: the kernel computes the DFT in float64 and only the final cast / log are float32, so
against the float64 oracle the bar is 1e-5 absolute in the log domain everywhere; against the
reference's output the bar is the reference's own complex64 FFT noise (see
tests/test_oracle.py::_specgram_noise_bound)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _noise_bound(ref_log):
    rel = ref_log.astype(np.float64) - ref_log.max(axis=1, keepdims=True)
    return 1e-5 + 3e-7 * np.exp(-rel / 2.0)


def test_log_specgram_matches_reference_golden(cuda_lib):
    from speech_b200.features import log_specgram_batch
    g = np.load(os.path.join(GOLD, "specgram.npz"))
    audios = [g["test0_audio"], g["test1_audio"]]
    feats, n_frames = log_specgram_batch(audios, 16000)
    feats = feats.cpu().numpy()
    assert feats.shape == (2, 156, 161) and n_frames == [109, 156]
    for e, name in enumerate(("test0", "test1")):
        ref = g[name + "_logspec"]
        got = feats[e, :n_frames[e]]
        assert (np.abs(got.astype(np.float64) - ref) <= _noise_bound(ref)).all()
        assert not feats[e, n_frames[e]:].any()          # zero padding past the utterance


@pytest.mark.parametrize("sr", [16000, 8000, 16050])        # 16050 Hz -> odd window (321)
def test_log_specgram_matches_float64_oracle_ragged_batch(cuda_lib, sr):
    from oracle.specgram_ref import log_specgram, preprocess
    from speech_b200.features import log_specgram_batch
    rng = np.random.RandomState(sr)
    nperseg = int(20 * sr / 1e3)
    lens = [sr, 3 * nperseg + 7, nperseg, nperseg - 1, 12345, 1]     # incl. 1 frame and 0 frames
    audios = [(rng.randn(n) * rng.choice([30, 3000, 20000])).clip(-32768, 32767).astype(np.int16)
              for n in lens]
    audios[4][100:4000] = 0                                           # digital silence: log(eps)
    feats, n_frames = log_specgram_batch(audios, sr)
    feats = feats.cpu().numpy()
    for e, a in enumerate(audios):
        ref = log_specgram(a, sr)
        assert n_frames[e] == ref.shape[0]
        assert np.abs(feats[e, :n_frames[e]] - ref).max(initial=0.0) < 1e-5
        assert not feats[e, n_frames[e]:].any()
    # normalised features (loader.py:65-67)
    nb = nperseg // 2 + 1
    mean = rng.randn(nb).astype(np.float32)
    std = (0.5 + rng.rand(nb)).astype(np.float32)
    featn, _ = log_specgram_batch(audios, sr, mean=mean, std=std)
    featn = featn.cpu().numpy()
    for e, a in enumerate(audios):
        ref = preprocess(a, sr, mean, std)
        assert np.abs(featn[e, :n_frames[e]] - ref).max(initial=0.0) < 5e-5


def test_featurised_batch_feeds_the_encoder(cuda_lib):
    """features -> CTC.loss: the (B, T, 161) tensor is what zero_pad_concat would have produced."""
    from speech_b200.loader import StagedBatch
    from speech_b200.features import log_specgram_batch
    from speech_b200.models import CTC
    rng = np.random.RandomState(0)
    audios = [(rng.randn(n) * 2000).astype(np.int16) for n in (16000, 12000, 14000)]
    labels = [[1, 2, 3], [4, 5], [6, 7, 8, 9]]
    cfg = {"dropout": 0.0, "encoder": {"conv": [[8, 5, 32, 2]],
                                       "rnn": {"dim": 32, "bidirectional": True, "layers": 1}}}
    torch.manual_seed(0)
    m = CTC(161, 10, cfg).cuda()
    feats, n_frames = log_specgram_batch(audios, 16000, mean=np.zeros(161, np.float32),
                                         std=np.full(161, 4.0, np.float32))
    host = [feats[e, :n].cpu().numpy() for e, n in enumerate(n_frames)]
    want = m.loss((host, labels)).item()
    x, y, x_lens, y_lens = m.collate(host, labels)
    got = m.loss(StagedBatch(feats, y, x_lens, y_lens, None)).item()
    assert got == want
