"""Golden teacher-forced outputs of the UNMODIFIED reference `Seq2Seq` (speech/models/seq2seq.py,
imported from /root/reference in this container) on a WSJ-shaped configuration (BASELINE.json
configs[3]: north-star conv stack + 3-layer biGRU-512 encoder, NNAttention with log_t, as
examples/wsj/seq2seq_config.json but with the BASELINE's width; dropout 0 and no scheduled
sampling so that the run is deterministic).  Stored: encoder states, logits, alignments, loss.
Weights are re-created from the seed on both sides; their checksum is stored.

    python tests/golden/make_golden_seq2seq_wsj.py    # rewrites tests/golden/seq2seq_wsj.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import load_ref_models  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FDIM, VOCAB, SEED = 80, 31, 1337
CFG = {"dropout": 0.0,
       "encoder": {"conv": [[32, 5, 8, 2], [32, 5, 8, 2]],
                   "rnn": {"dim": 512, "bidirectional": True, "layers": 3}},
       "decoder": {"embedding_dim": 512, "layers": 1, "log_t": True, "sample_prob": 0}}


def batch():
    rng = np.random.RandomState(SEED)
    inputs = [rng.randn(n, FDIM).astype(np.float32) for n in (100, 92)]
    # <s> = VOCAB-1 first, </s> = VOCAB-2 last (Preprocessor.encode with start_and_end)
    labels = [[VOCAB - 1] + rng.randint(0, VOCAB - 2, n).tolist() + [VOCAB - 2] for n in (7, 5)]
    return inputs, labels


def weight_checksum(model):
    return float(sum(p.detach().double().abs().sum().item() for p in model.parameters()))


def main():
    Seq2Seq = load_ref_models()["seq2seq"].Seq2Seq
    torch.manual_seed(SEED)
    m = Seq2Seq(FDIM, VOCAB, CFG)
    m.eval()
    b = batch()
    with torch.no_grad():
        x, y = m.collate(*b)
        x_enc = m.encode(x)
        out, aligns = m.decode(x_enc, y)
        loss = m.loss(b)
    g = {"x_enc": x_enc.numpy(), "out": out.numpy(), "aligns": aligns.numpy(),
         "loss": np.float64(float(loss)), "wsum": np.float64(weight_checksum(m))}
    np.savez_compressed(os.path.join(OUT, "seq2seq_wsj.npz"), **g)
    for k, v in g.items():
        print(k, getattr(v, "shape", ()), float(v) if np.ndim(v) == 0 else "")


if __name__ == "__main__":
    main()
