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
