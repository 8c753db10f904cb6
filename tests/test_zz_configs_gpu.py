"""GPU parity on the encoder configurations BASELINE.json names, against outputs of the UNMODIFIED
reference `Model.encode` (speech/models/model.py:60-79) generated in the build container by
tests/golden/make_golden_configs.py: the shipped TIMIT CTC recipe (4-layer biGRU-256 on 161 bins,
second conv layer 5x32 stride 1), a 4-layer biGRU-512 and the north-star 5-layer biGRU-1024.

Weights are not stored: both sides build the model under the same torch seed (the drop-in creates
its parameters in the reference's order) and the fixture's weight checksum is asserted first.
Tolerance: bf16 tensor-core operands against the reference's fp32 CPU arithmetic.  Rounding the
GEMM / recurrent operands of the reference model to bf16 on the CPU (same places as the kernels)
moves these outputs by at most 2.1e-3 (mean 4.3e-4) on the three fixtures; the bars are 1e-2 on
every element and 2e-3 on average."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)


@pytest.mark.parametrize("tag", ["timit", "timit512", "libri"])
def test_encode_matches_reference_on_baseline_configs(cuda_lib, tag):
    from make_golden_configs import CONFIGS, weight_checksum
    from speech_b200.models import Model
    g = np.load(os.path.join(GOLD, "encoder_configs.npz"))
    fdim, cfg, seed, B, T = CONFIGS[tag]
    torch.manual_seed(seed)
    m = Model(fdim, cfg)
    assert abs(weight_checksum(m) - float(g[tag + "_wsum"])) <= 1e-9 * float(g[tag + "_wsum"])
    m.cuda()
    m.set_eval()
    with torch.no_grad():
        y = m.encode(torch.from_numpy(g[tag + "_x"]).cuda())
    ref = g[tag + "_y"]
    assert tuple(y.shape) == ref.shape == (B, m.conv_out_size(T, 0), cfg["encoder"]["rnn"]["dim"])
    err = np.abs(y.float().cpu().numpy() - ref)
    assert err.max() < 1e-2, err.max()
    assert err.mean() < 2e-3, err.mean()


@pytest.mark.parametrize("tag", ["tiny", "timit"])
def test_transducer_lattice_matches_reference_output(cuda_lib, tag):
    """`Transducer.forward_impl` = decode(encode(x), y) -> (B, T', U+1, V+1) log-probabilities
    (transducer_model.py:38-77) against the reference's own output (make_golden_transducer.py),
    same seed => same weights.  The encoder and the prediction network run on the GRU kernels
    (bf16 operands), the joint in fp32: emulating the operand rounding on the CPU moves the
    lattice by <= 9e-4 (mean 1.7e-4); bars 1e-2 / 1.5e-3."""
    from make_golden_transducer import CONFIGS as TCONF, batch_for, weight_checksum
    from speech_b200.models import Transducer
    g = np.load(os.path.join(GOLD, "transducer.npz"))
    fdim, vocab, cfg, seed, _, _ = TCONF[tag]
    torch.manual_seed(seed)
    m = Transducer(fdim, vocab, cfg)
    assert abs(weight_checksum(m) - float(g[tag + "_wsum"])) <= 1e-9 * float(g[tag + "_wsum"])
    m.cuda()
    m.set_eval()
    out = m(batch_for(tag))
    ref = g[tag + "_out"]
    assert tuple(out.shape) == ref.shape
    err = np.abs(out.float().cpu().numpy() - ref)
    assert err.max() < 1e-2, err.max()
    assert err.mean() < 1.5e-3, err.mean()
    # rows are normalised log-probabilities
    assert np.allclose(np.exp(out.double().cpu().numpy()).sum(-1), 1.0, atol=1e-5)


def test_seq2seq_wsj_shape_matches_reference_output(cuda_lib):
    """WSJ-shaped attention model (BASELINE.json configs[3]: north-star conv stack, 3-layer
    biGRU-512, NNAttention with log_t): encoder states, teacher-forced logits, alignments and the
    loss of the reference's own run (make_golden_seq2seq_wsj.py), same seed => same weights.
    Emulating the kernels' bf16 operand rounding on the CPU moves them by 1.3e-3 / 3.2e-4 / 3e-5 /
    5e-6 (relative); bars 1e-2 / 5e-3 / 1e-3 / 1e-4."""
    from make_golden_seq2seq_wsj import CFG, FDIM, SEED, VOCAB, batch, weight_checksum
    from speech_b200.models import Seq2Seq
    g = np.load(os.path.join(GOLD, "seq2seq_wsj.npz"))
    torch.manual_seed(SEED)
    m = Seq2Seq(FDIM, VOCAB, CFG)
    assert abs(weight_checksum(m) - float(g["wsum"])) <= 1e-9 * float(g["wsum"])
    m.cuda()
    m.set_eval()
    b = batch()
    with torch.no_grad():
        x, y = m.collate(*b)
        x_enc = m.encode(x.cuda())
        out, aligns = m.decode(x_enc, y.cuda())
        loss = m.loss(b)
    assert np.abs(x_enc.cpu().numpy() - g["x_enc"]).max() < 1e-2
    assert np.abs(out.cpu().numpy() - g["out"]).max() < 5e-3
    assert np.abs(aligns.cpu().numpy() - g["aligns"]).max() < 1e-3
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < 1e-4
