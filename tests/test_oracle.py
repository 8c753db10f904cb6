"""CPU tests (-m "not gpu"): the oracle restatements against the reference's golden vectors and
against independent implementations; host-side logic; the C-ABI library's exported symbols."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import ctc_decoder_ref, ctc_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decoder_oracle_matches_reference_golden_vectors():
    g = np.load(os.path.join(GOLD, "ctc_decode.npz"))
    cases = [("demo", 0, (10, 8, 1)), ("f32", 11, (1, 4, 10)), ("peaky", 4, (1, 3, 8)),
             ("tie", 0, (1, 3))]
    for name, blank, beams in cases:
        probs = g[name + "_probs"]
        if name == "f32":
            probs = probs.astype(np.float64)
        for b in beams:
            lab, sc = ctc_decoder_ref.prefix_beam_search(probs, beam_size=b, blank=blank)
            assert list(lab) == list(g["%s_b%d_labels" % (name, b)]), (name, b)
            assert abs(sc - float(g["%s_b%d_score" % (name, b)])) < 1e-9


def test_decoder_demo_matches_survey_captured_values():
    # the reference file's own __main__ demo, captured during the survey
    g = np.load(os.path.join(GOLD, "ctc_decode.npz"))
    assert list(g["demo_b10_labels"]) == [5, 12, 8, 2, 16, 13, 3, 7, 8, 3, 10, 4, 11, 19, 10, 14, 9,
                                          8, 6, 5, 19, 4, 5, 4, 6]
    assert abs(float(g["demo_b10_score"]) - 100.601941) < 1e-5
    assert abs(float(g["demo_b8_score"]) - 102.401333) < 1e-5
    assert abs(float(g["demo_b1_score"]) - 112.415479) < 1e-5


@pytest.mark.parametrize("B,T,V,seed", [(3, 20, 7, 0), (4, 48, 11, 1), (2, 35, 29, 2)])
def test_ctc_oracle_matches_torch_ctc_loss(B, T, V, seed):
    rng = np.random.RandomState(seed)
    acts = rng.randn(B, T, V).astype(np.float32)
    llen = rng.randint(0, 9, size=B).astype(np.int32)
    flat = np.concatenate([rng.randint(0, V - 1, size=L) for L in llen] + [np.zeros(0, int)]).astype(np.int32)
    alen = rng.randint(T // 2 + 9, T + 1, size=B).astype(np.int32)
    c, g = ctc_ref.ctc_loss_and_grad(acts, flat, alen, llen)
    a = torch.from_numpy(acts).double().requires_grad_(True)
    lp = torch.log_softmax(a, 2).transpose(0, 1)
    loss = torch.nn.functional.ctc_loss(lp, torch.from_numpy(flat).long(), torch.from_numpy(alen).long(),
                                        torch.from_numpy(llen).long(), blank=V - 1, reduction="none")
    loss.sum().backward()
    np.testing.assert_allclose(c, loss.detach().numpy(), rtol=1e-10)
    assert np.abs(g - a.grad.numpy()).max() < 1e-10


def test_max_decode_known_answers():
    # reference tests/ctc_test.py:31-43
    from speech_b200.models import CTC
    assert CTC.max_decode([1, 2, 2, 0, 0, 0, 2, 1], 0) == [1, 2, 2, 1]
    assert CTC.max_decode([2, 2, 2], 0) == [2]
    assert CTC.max_decode([0, 0, 0], 0) == []


def test_model_class_surface_and_state_dict_names():
    from speech_b200.models import CTC, Model
    cfg = {"dropout": 0.0, "encoder": {"conv": [[32, 5, 32, 2]],
                                       "rnn": {"dim": 16, "bidirectional": False, "layers": 1}}}
    m = Model(40, cfg)
    assert m.conv_out_size(100, 0) == 48 and m.conv_out_size(40, 1) == 5
    assert m.encoder_dim == 16 and not m.is_cuda
    g = np.load(os.path.join(GOLD, "encoder_tiny.npz"))
    ref_keys = {k[len("tiny__sd__"):].replace("__", ".") for k in g.files if k.startswith("tiny__sd__")}
    assert set(m.state_dict().keys()) == ref_keys
    c = CTC(40, 10, cfg)
    assert c.blank == 10 and c.fc.fc.weight.shape == (11, 16)
    x, y, xl, yl = c.collate([np.zeros((100, 40)), np.zeros((90, 40))], [[1, 2, 3], [4]])
    assert x.shape == (2, 100, 40) and x.dtype == torch.float32
    assert xl.tolist() == [48, 48] and yl.tolist() == [3, 1] and y.tolist() == [1, 2, 3, 4]
    with pytest.raises(Exception):
        c.loss(([np.zeros((100, 40))], [[1]]))     # CPU model: must fail loudly, no fallback


def test_same_seed_same_init_as_reference():
    """identical construction order => identical weights under the same torch seed."""
    from speech_b200.models import Model
    cfg = {"dropout": 0.0, "encoder": {"conv": [[32, 5, 32, 2]],
                                       "rnn": {"dim": 16, "bidirectional": False, "layers": 1}}}
    torch.manual_seed(0)
    m = Model(40, cfg)
    g = np.load(os.path.join(GOLD, "encoder_tiny.npz"))
    for k, v in m.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), g["tiny__sd__" + k.replace(".", "__")])


def test_library_exports_every_declared_symbol():
    from speech_b200.csrc import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    hdr = open(os.path.join(ROOT, "include", "speech_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), n
    from speech_b200 import _lib
    assert set(_lib.SIGNATURES) == names
    assert lib.sb_version() >= 100


def test_workspace_size_queries_need_no_gpu():
    """The size queries of the C ABI are pure host arithmetic: they answer on a CPU box, grow with
    the problem and reject non-positive sizes (-1 = SB_ERR_INVALID)."""
    from speech_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    assert lib.sb_s2s_workspace_size(16, 200, 512, ctypes.byref(n)) == 0
    small = n.value
    # tickets + scores (B*T) + 4 scalars and 2 H-wide partials per CTA of 24 frames
    ts = (200 + 23) // 24
    assert small >= 4 * (16 * 200 + 4 * 16 * ts + 2 * 16 * ts * 512)
    assert lib.sb_s2s_workspace_size(16, 400, 512, ctypes.byref(n)) == 0 and n.value > small
    assert lib.sb_s2s_workspace_size(0, 200, 512, ctypes.byref(n)) != 0
    assert lib.sb_gru_fwd_workspace_size(64, 1024, 2, ctypes.byref(n)) == 0 and n.value >= 1024
    assert lib.sb_gru_bwd_workspace_size(64, 1024, 2, ctypes.byref(n)) == 0
    assert n.value >= 1024 + 2 * 2 * 64 * 3 * 1024 * 2          # counters + bf16 exchange tiles
    assert lib.sb_gru_fwd_workspace_size(64, 1024, 3, ctypes.byref(n)) != 0


def test_c_oracle_matches_numpy_oracle():
    from oracle import build as ob
    rng = np.random.RandomState(4)
    B, T, V = 5, 40, 9
    acts = rng.randn(B, T, V).astype(np.float32)
    llen = np.array([0, 3, 7, 12, 19], np.int32)
    flat = np.concatenate([rng.randint(0, 3, size=L) for L in llen]).astype(np.int32)
    alen = np.array([40, 33, 40, 25, 38], np.int32)
    c1, g1 = ctc_ref.ctc_loss_and_grad(acts, flat, alen, llen)
    c2, g2 = ob.ctc(acts, flat, llen, alen, V - 1)
    np.testing.assert_allclose(c2, c1, rtol=1e-12)
    assert np.abs(g2 - g1).max() < 1e-6


def test_rnnt_oracle_matches_brute_force_enumeration():
    from oracle import rnnt_ref
    rng = np.random.RandomState(0)
    for T, U, V in [(1, 0, 3), (3, 2, 4), (4, 3, 3), (5, 1, 5)]:
        x = rng.randn(T, U + 1, V)
        lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
        labels = [int(v) for v in rng.randint(0, V - 1, size=U)]
        c, g = rnnt_ref.rnnt_single(lp, labels, V - 1)
        assert abs(c - rnnt_ref.brute_force_nll(lp, labels, V - 1)) < 1e-10
        # gradient vs central finite differences of the brute-force definition
        eps = 1e-6
        for idx in [(0, 0, V - 1), (T - 1, U, V - 1)] + ([(0, 0, labels[0])] if U else []):
            lp2 = lp.copy(); lp2[idx] += eps
            lp3 = lp.copy(); lp3[idx] -= eps
            fd = (rnnt_ref.brute_force_nll(lp2, labels, V - 1) -
                  rnnt_ref.brute_force_nll(lp3, labels, V - 1)) / (2 * eps)
            assert abs(fd - g[idx]) < 1e-6


def _specgram_noise_bound(ref_log):
    """The reference computes its FFT in complex64 (scipy: result_type(int16, complex64)), so its
    own output carries rounding noise ~ eps32 * sqrt(N) * (frame maximum / bin amplitude) relative
    to exact arithmetic.  In the log-power domain that is 3e-7 * exp(-rel/2) for a bin whose log
    power lies `rel` below the frame maximum (measured on the fixtures: 1.2e-6 at rel > -5,
    1.5e-4 at -15, 1.4e-3 at -20), plus 1e-5 for the float32 log itself."""
    rel = ref_log.astype(np.float64) - ref_log.max(axis=1, keepdims=True)
    return 1e-5 + 3e-7 * np.exp(-rel / 2.0)


def test_specgram_oracle_matches_reference_output_on_its_own_fixtures():
    """oracle/specgram_ref.py (float64 restatement of loader.py:156-166) against the output of
    the reference's own function on tests/test0.wav / test1.wav (golden/make_specgram_golden.py)."""
    from oracle.specgram_ref import log_specgram
    g = np.load(os.path.join(GOLD, "specgram.npz"))
    for name in ("test0", "test1"):
        ours = log_specgram(g[name + "_audio"], int(g[name + "_sr"]))
        ref = g[name + "_logspec"]
        assert ours.shape == ref.shape and ours.dtype == np.float32
        assert (np.abs(ours.astype(np.float64) - ref) <= _specgram_noise_bound(ref)).all()
    # loader_test.py:19-20 pins (time, freq) orientation and float32; wave_test.py:16 the duration
    assert g["test0_logspec"].shape[1] == 161 and round(g["test0_audio"].shape[0] / 16000, 3) == 1.101


def test_specgram_oracle_matches_scipy_on_random_audio():
    import scipy.signal
    from oracle.specgram_ref import log_specgram
    rng = np.random.RandomState(0)
    for sr, n in ((16000, 16000), (8000, 5000), (16050, 4001)):     # 16050 Hz: odd nperseg (321)
        audio = (rng.randn(n) * 3000).astype(np.int16)
        nperseg, noverlap = int(20 * sr / 1e3), int(10 * sr / 1e3)
        _, _, spec = scipy.signal.spectrogram(audio, fs=sr, window="hann", nperseg=nperseg,
                                              noverlap=noverlap, detrend=False)
        ref = np.log(spec.T.astype(np.float32) + 1e-10)
        ours = log_specgram(audio, sr)
        assert ours.shape == ref.shape
        assert (np.abs(ours.astype(np.float64) - ref) <= _specgram_noise_bound(ref)).all()
