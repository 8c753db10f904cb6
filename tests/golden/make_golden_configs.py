"""Golden encoder outputs of the UNMODIFIED reference (`Model.encode`, speech/models/model.py:60-79,
imported from /root/reference in this container) on the encoder configurations BASELINE.json
names, at sizes the CPU finishes in seconds:

  timit       the shipped TIMIT CTC recipe (examples/timit/ctc_config.json): conv
              [[32,5,32,2],[32,5,32,1]], 4-layer biGRU-256, 161 log-spectrogram bins
  timit512    BASELINE.json configs[1]: 4-layer biGRU-512 on 80 features (north-star conv stack)
  libri       BASELINE.json configs[2] = the north-star encoder: 5-layer biGRU-1024, 80 features

The weights are NOT stored (85 M parameters for `libri`): both sides construct the model under
`torch.manual_seed(seed)`; the drop-in creates its parameters in the reference's order
(tests/test_oracle.py::test_same_seed_same_init_as_reference), and the fixture carries a checksum
of the reference's weights so that the GPU test can assert it is comparing like with like.
Dropout is 0 (parity runs).

    python tests/golden/make_golden_configs.py        # rewrites tests/golden/encoder_configs.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import load_ref_models  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NS_CONV = [[32, 5, 8, 2], [32, 5, 8, 2]]
CONFIGS = {
    # tag: (feature dim, cfg, seed, B, T)
    "timit": (161, {"dropout": 0.0, "encoder": {"conv": [[32, 5, 32, 2], [32, 5, 32, 1]],
                                                "rnn": {"dim": 256, "bidirectional": True,
                                                        "layers": 4}}}, 2017, 3, 90),
    "timit512": (80, {"dropout": 0.0, "encoder": {"conv": NS_CONV,
                                                  "rnn": {"dim": 512, "bidirectional": True,
                                                          "layers": 4}}}, 11, 4, 120),
    "libri": (80, {"dropout": 0.0, "encoder": {"conv": NS_CONV,
                                               "rnn": {"dim": 1024, "bidirectional": True,
                                                       "layers": 5}}}, 0, 2, 100),
}


def weight_checksum(model):
    return float(sum(p.detach().double().abs().sum().item() for p in model.parameters()))


def main():
    Model = load_ref_models()["model"].Model
    torch.set_num_threads(8)
    g = {}
    for tag, (fdim, cfg, seed, B, T) in CONFIGS.items():
        torch.manual_seed(seed)
        m = Model(fdim, cfg)
        x = torch.randn(B, T, fdim)
        with torch.no_grad():
            y = m.encode(x)
        g[tag + "_x"] = x.numpy()
        g[tag + "_y"] = y.numpy()
        g[tag + "_wsum"] = np.float64(weight_checksum(m))
        print(tag, tuple(x.shape), "->", tuple(y.shape), "wsum %.6f" % g[tag + "_wsum"])
    np.savez_compressed(os.path.join(OUT, "encoder_configs.npz"), **g)


if __name__ == "__main__":
    main()
