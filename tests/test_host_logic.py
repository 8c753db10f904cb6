"""CPU tests of the host-side mirror of the reference interface (SURVEY.md section 8 rows a2, a5,
a19, b): batch assembly conventions, the conv output-size rule, and that the product path fails
loudly - never falls back to a CPU implementation - when there is no CUDA device."""
import numpy as np
import pytest
import torch

TINY = {"dropout": 0.0, "encoder": {"conv": [[32, 5, 32, 2]],
                                    "rnn": {"dim": 16, "bidirectional": False, "layers": 1}},
        "decoder": {"embedding_dim": 16, "layers": 2}}
WSJ = {"dropout": 0.4, "encoder": {"conv": [[32, 5, 8, 2], [32, 5, 8, 2]],
                                   "rnn": {"dim": 32, "bidirectional": True, "layers": 2}}}


def _batch(n=4, seed=0):
    rng = np.random.RandomState(seed)
    inputs = [rng.randn(50 + 9 * e, 40) for e in range(n)]          # float64, ragged lengths
    labels = [rng.randint(0, 10, 4 + e).tolist() for e in range(n)]
    return inputs, labels


def test_zero_pad_concat_pads_with_zeros_and_casts_to_float32():
    """reference model.py:135-141: (B, max T, F) float32, zero padded; fixtures arrive float64."""
    from speech_b200.models.model import zero_pad_concat
    inputs, _ = _batch()
    cat = zero_pad_concat(inputs)
    assert cat.dtype == np.float32 and cat.shape == (4, 77, 40)
    for e, inp in enumerate(inputs):
        assert np.array_equal(cat[e, :inp.shape[0]], inp.astype(np.float32))
        assert not cat[e, inp.shape[0]:].any()


@pytest.mark.parametrize("cfg,fdim", [(TINY, 40), (WSJ, 80)])
def test_conv_out_size_equals_torch_valid_convolution(cfg, fdim):
    """reference model.py:44-52: ceil((n-k+1)/s) per layer == torch's floor((n-k)/s)+1."""
    from speech_b200.models import CTC
    m = CTC(fdim, 10, cfg)
    for n in (57, 100, 101, 333, 1000):
        x = torch.zeros(1, 1, n, fdim)
        with torch.no_grad():
            for c in m.conv.children():
                if isinstance(c, torch.nn.Conv2d):
                    x = c(x)
        assert m.conv_out_size(n, 0) == x.shape[2]
        assert m.conv_out_size(fdim, 1) == x.shape[3]
    assert m.rnn.input_size == x.shape[1] * x.shape[3]


def test_dropout_shifts_conv_state_dict_indices_like_the_reference():
    """model.py:25-26: a Dropout module after every ReLU => conv.{0,3,..} instead of conv.{0,2,..}."""
    from speech_b200.models import CTC
    keys = set(CTC(80, 10, WSJ).state_dict().keys())
    assert {"conv.0.weight", "conv.3.weight"} <= keys and "conv.2.weight" not in keys
    nodrop = dict(WSJ, dropout=0.0)
    keys = set(CTC(80, 10, nodrop).state_dict().keys())
    assert {"conv.0.weight", "conv.2.weight"} <= keys


def test_ctc_collate_conventions():
    """reference ctc_model.py:42-53: x_lens = T' of the padded batch for EVERY utterance, labels
    flat int32 on the host, y_lens int32."""
    from speech_b200.models import CTC
    m = CTC(40, 10, TINY)
    inputs, labels = _batch()
    x, y, x_lens, y_lens = m.collate(inputs, labels)
    assert x.shape == (4, 77, 40) and x.dtype == torch.float32
    assert x_lens.dtype == torch.int32 and x_lens.tolist() == [m.conv_out_size(77, 0)] * 4
    assert y.dtype == torch.int32 and y.tolist() == [t for l in labels for t in l]
    assert y_lens.tolist() == [len(l) for l in labels]


def test_seq2seq_and_transducer_label_padding_uses_the_end_token():
    """reference seq2seq.py:239-248 / transducer_model.py:103-116: labels are padded with
    labels[0][-1] (the end token), int64."""
    from speech_b200.models import Seq2Seq, Transducer
    from speech_b200.models.seq2seq import end_pad_concat
    labels = [[7, 1, 2, 9], [7, 3, 9], [7, 4, 5, 6, 8, 9]]
    cat = end_pad_concat(labels)
    assert cat.dtype == np.int64 and cat.shape == (3, 6)
    assert cat[1].tolist() == [7, 3, 9, 9, 9, 9] and cat[0].tolist() == [7, 1, 2, 9, 9, 9]
    s = Seq2Seq(40, 10, TINY)
    inputs = [np.zeros((60, 40)), np.zeros((50, 40)), np.zeros((55, 40))]
    x, y = s.collate(inputs, labels)
    assert x.shape == (3, 60, 40) and torch.equal(y, torch.from_numpy(cat))
    t = Transducer(40, 10, TINY)
    assert torch.equal(t.label_collate(labels), torch.from_numpy(cat))
    x, yf, x_lens, y_lens = t.collate(inputs, labels)
    assert yf.dtype == torch.int32 and y_lens.tolist() == [4, 3, 6]
    assert x_lens.tolist() == [t.conv_out_size(60, 0)] * 3


def test_product_path_raises_without_cuda_instead_of_falling_back():
    """No CPU implementation is shipped: every entry point that would have to compute must raise
    SpeechB200Error on CPU tensors (the oracle is test infrastructure only)."""
    from speech_b200 import _lib
    from speech_b200.functions.ctc import CTCLoss
    from speech_b200.models import CTC
    from speech_b200.models.ctc_decoder import decode
    m = CTC(40, 10, TINY)          # parameters on the CPU
    with pytest.raises(_lib.SpeechB200Error):
        m.loss(_batch())
    with pytest.raises(_lib.SpeechB200Error):
        m.infer(_batch())
    acts = torch.randn(2, 20, 11, requires_grad=True)
    with pytest.raises(_lib.SpeechB200Error):
        CTCLoss()(acts, torch.IntTensor([1, 2, 3]), torch.IntTensor([20, 20]),
                  torch.IntTensor([2, 1]))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SpeechB200Error):
            decode(np.full((5, 4), 0.25, dtype=np.float32), beam_size=2, blank=0)


def test_batch_prefetcher_requires_a_cuda_model():
    from speech_b200.loader import BatchPrefetcher
    from speech_b200.models import CTC
    with pytest.raises(RuntimeError):
        BatchPrefetcher(CTC(40, 10, TINY), [])


def test_shims_resolve_to_the_library_operators():
    """INTEGRATION.md option A: `import functions.ctc`, `import transducer...` from shims/."""
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "shims"))
    try:
        fc = importlib.import_module("functions.ctc")
        tf = importlib.import_module("transducer.functions.transducer")
        td = importlib.import_module("transducer.decoders")
    finally:
        sys.path.pop(0)
    from speech_b200.functions import ctc, transducer
    assert fc.CTCLoss is ctc.CTCLoss
    assert tf.TransducerLoss is transducer.TransducerLoss
    assert callable(td.decode_static)


def _levenshtein(a, b):
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def test_edit_distance_and_cer_match_the_definition():
    """speech/utils/score.py:7-18 with `editdistance.eval` replaced by the library's host function
    sb_edit_distance: known answers plus a plain-Python dynamic programme on random sequences."""
    import speech_b200
    from speech_b200.utils.score import compute_cer, edit_distance
    assert edit_distance("kitten", "sitting") == 3
    assert edit_distance("", "abc") == 3 and edit_distance("abc", "") == 3
    assert edit_distance([], []) == 0
    assert edit_distance(("sil", "ae", "t"), ("sil", "t")) == 1          # phoneme tokens
    rng = np.random.RandomState(0)
    results = []
    for _ in range(50):
        a = rng.randint(0, 6, rng.randint(0, 40)).tolist()
        b = rng.randint(0, 6, rng.randint(0, 40)).tolist()
        assert edit_distance(a, b) == _levenshtein(a, b) == edit_distance(b, a)
        results.append((a, b))
    results = [r for r in results if len(r[0])]
    want = sum(_levenshtein(a, b) for a, b in results) / sum(len(a) for a, _ in results)
    assert compute_cer(results) == want
    assert speech_b200.compute_cer is compute_cer                         # speech/__init__.py:2


def test_save_load_round_trip_like_the_reference_io_test(tmp_path):
    """tests/io_test.py of the reference: whole-module pickle + preprocessor pickle under the same
    file names; plus the resume state the reference lacks."""
    import pickle
    import speech_b200
    from speech_b200.models import CTC
    from speech_b200.utils import io

    class Preproc:                         # stands in for speech.loader.Preprocessor
        def __init__(self):
            self.mean, self.std = np.zeros(3), np.ones(3)
            self.int_to_char = {0: "a"}
            self.char_to_int = {"a": 0}
    globals()["Preproc"] = Preproc
    Preproc.__qualname__ = "Preproc"
    Preproc.__module__ = __name__
    torch.manual_seed(0)
    m = CTC(40, 10, TINY)
    speech_b200.save(m, Preproc(), str(tmp_path))
    assert sorted(p.name for p in tmp_path.iterdir()) == ["model", "preproc.pyc"]
    speech_b200.save(m, Preproc(), str(tmp_path), tag="best")
    assert (tmp_path / "best_model").exists() and (tmp_path / "best_preproc.pyc").exists()
    s_model, s_preproc = speech_b200.load(str(tmp_path))
    for attr in ("mean", "std", "int_to_char", "char_to_int"):
        assert hasattr(s_preproc, attr)
    msd = m.state_dict()
    for k, v in s_model.state_dict().items():
        assert k in msd and torch.equal(v, msd[k])
    assert hasattr(s_model, "encoder_dim") and hasattr(s_model, "is_cuda")
    # resume state: state_dict + optimiser + counters, written atomically
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9)
    m.fc.fc.weight.grad = torch.ones_like(m.fc.fc.weight)
    opt.step()
    name = io.save_state(m, str(tmp_path), optimizer=opt, epoch=3, iteration=1234)
    assert name.endswith("state") and not (tmp_path / "state.tmp").exists()
    torch.manual_seed(1)
    m2 = CTC(40, 10, TINY)
    opt2 = torch.optim.SGD(m2.parameters(), lr=1e-3, momentum=0.9)
    counters = io.load_state(m2, str(tmp_path), optimizer=opt2)
    assert counters == {"epoch": 3, "iteration": 1234}
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert len(opt2.state_dict()["state"]) == len(opt.state_dict()["state"])


def test_stage_inputs_routes_to_the_device_assembler_when_the_model_is_on_a_gpu(monkeypatch):
    """The CUDA branch of Model.stage_inputs cannot run here; check its wiring with a stand-in."""
    from speech_b200.models import CTC, model as model_mod
    m = CTC(40, 10, TINY)
    calls = []

    def fake(inputs, device):
        calls.append((len(inputs), device))
        return torch.from_numpy(model_mod.zero_pad_concat(inputs))
    monkeypatch.setattr(model_mod, "zero_pad_concat_device", fake)
    monkeypatch.setattr(type(m), "is_cuda", property(lambda self: True))
    inputs, labels = _batch()
    x, y, x_lens, y_lens = m.collate(inputs, labels)
    assert calls == [(4, next(m.parameters()).device)]
    assert x.shape == (4, 77, 40) and y.dtype == torch.int32


def test_conv_stack_has_no_library_fallback():
    """Every stack the reference can express runs on the package's own kernels, training included
    (the TIMIT recipes' second layer [*, 5, 32, 1] used to fall back to nn.Conv2d for backward):
    on a CPU tensor conv_stack must therefore raise instead of silently taking an nn-module route,
    and a stack the kernels do not cover (padding) must raise too."""
    from speech_b200 import _lib, ops
    from speech_b200.models import CTC
    timit = {"dropout": 0.0, "encoder": {"conv": [[8, 5, 32, 2], [8, 5, 32, 1]],
                                         "rnn": {"dim": 16, "bidirectional": True, "layers": 1}}}
    m = CTC(161, 10, timit)
    x = torch.randn(2, 40, 161)
    with pytest.raises(_lib.SpeechB200Error):
        ops.conv_stack(x, m.conv, True)
    wsj = CTC(80, 10, WSJ)
    with pytest.raises(_lib.SpeechB200Error):
        ops.conv_stack(torch.randn(2, 40, 80), wsj.conv, True)
    with torch.no_grad(), pytest.raises(_lib.SpeechB200Error):
        ops.conv_stack(x, m.conv, False)
    padded = torch.nn.Sequential(torch.nn.Conv2d(1, 8, (5, 8), stride=(2, 2), padding=1),
                                 torch.nn.ReLU())
    with pytest.raises(_lib.SpeechB200Error):
        ops.conv_stack(torch.randn(2, 40, 80), padded, False)
