"""Golden RNN-T lattices of the UNMODIFIED reference `Transducer.forward_impl` =
`decode(encode(x), y)` (speech/models/transducer_model.py:38-77, imported from /root/reference in
this container; the un-vendored `transducer` package it imports at module level is stubbed - the
forward pass never touches it).  Two configurations:

  tiny    conv [[8,5,32,2]], 2-layer biGRU-32 encoder, 1-layer GRU prediction network
  timit   the shipped recipe examples/timit/transducer_config.json (conv [[8,5,32,2],[8,5,32,1]],
          4-layer biGRU-256, embedding 256), dropout 0 for parity

Weights are not stored: both sides build the model under `torch.manual_seed(seed)` (the drop-in
creates its parameters in the reference's order); a checksum of the reference's weights is kept.

    python tests/golden/make_golden_transducer.py     # rewrites tests/golden/transducer.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, load_ref_models  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CONFIGS = {
    # tag: (feature dim, vocab, cfg, seed, input lengths, label lengths)
    "tiny": (40, 10, {"dropout": 0.0,
                      "encoder": {"conv": [[8, 5, 32, 2]],
                                  "rnn": {"dim": 32, "bidirectional": True, "layers": 2}},
                      "decoder": {"embedding_dim": 32, "layers": 1}}, 0, (60, 52, 57), (5, 7, 3)),
    "timit": (161, 50, {"dropout": 0.0,
                        "encoder": {"conv": [[8, 5, 32, 2], [8, 5, 32, 1]],
                                    "rnn": {"dim": 256, "bidirectional": True, "layers": 4}},
                        "decoder": {"embedding_dim": 256, "layers": 1}}, 2017, (70, 64), (6, 4)),
}


def batch_for(tag):
    fdim, vocab, cfg, seed, in_lens, lab_lens = CONFIGS[tag]
    rng = np.random.RandomState(seed + 1)
    inputs = [rng.randn(n, fdim).astype(np.float32) for n in in_lens]
    labels = [rng.randint(0, vocab, n).tolist() for n in lab_lens]
    return inputs, labels


def weight_checksum(model):
    return float(sum(p.detach().double().abs().sum().item() for p in model.parameters()))


def load_ref_transducer():
    load_ref_models()                                   # registers refmodels.model
    for name in ("transducer", "transducer.decoders", "transducer.functions",
                 "transducer.functions.transducer"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location(
        "refmodels.transducer_model", os.path.join(REF, "speech", "models", "transducer_model.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refmodels.transducer_model"] = mod
    spec.loader.exec_module(mod)
    return mod.Transducer


def main():
    Transducer = load_ref_transducer()
    g = {}
    for tag, (fdim, vocab, cfg, seed, _, _) in CONFIGS.items():
        torch.manual_seed(seed)
        m = Transducer(fdim, vocab, cfg)
        m.eval()
        inputs, labels = batch_for(tag)
        x, y, x_lens, y_lens = m.collate(inputs, labels)
        y_mat = m.label_collate(labels)
        with torch.no_grad():
            out = m.forward_impl(x, y_mat)
        g[tag + "_out"] = out.numpy()
        g[tag + "_wsum"] = np.float64(weight_checksum(m))
        print(tag, tuple(x.shape), "->", tuple(out.shape), "wsum %.6f" % g[tag + "_wsum"])
    np.savez_compressed(os.path.join(OUT, "transducer.npz"), **g)


if __name__ == "__main__":
    main()
