"""Generate golden vectors by running the UNMODIFIED reference (awni/speech at /root/reference)
in this container.  The reference cannot travel to the GPU box, so its inputs/outputs are frozen
here as small .npz fixtures (committed) that both the CPU oracle tests and the GPU parity tests read.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

What is pinned:
  ctc_decode.npz     ctc_decoder.decode (speech/models/ctc_decoder.py:38-113): the file's own demo
                     (seed 3, T=50, S=20, blank 0; beams 10/8/1) + float32 / blank-last / repeat cases
  encoder_tiny.npz   Model.encode (speech/models/model.py:60-79) on tests/shared.py's tiny config and
                     on a 2-layer bidirectional config, weights + input + output
  ctc_tiny.npz       CTC.forward_impl logits (speech/models/ctc_model.py:25-32) for the tiny config
  seq2seq_tiny.npz   Seq2Seq teacher-forced forward / decode_step / infer / beam_search
                     (speech/models/seq2seq.py) with seeds 1337 as tests/seq2seq_test.py:14-15.
                     beam_search needs the py3 one-token fix list(filter(...)) (seq2seq.py:211),
                     applied IN MEMORY to the source text before exec; nothing is copied.
The CTC loss / transducer loss live in un-vendored dependencies (Makefile:4-12) and cannot be run.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_ref_models():
    """Import speech/models/{model,ctc_decoder,seq2seq}.py without speech/__init__ (which needs
    the absent warp-ctc / transducer / editdistance packages)."""
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [os.path.join(REF, "speech", "models")]
    sys.modules["refmodels"] = pkg
    mods = {}
    for name in ("model", "ctc_decoder"):
        spec = importlib.util.spec_from_file_location(
            "refmodels." + name, os.path.join(REF, "speech", "models", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["refmodels." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
        setattr(pkg, name, m)
    src = open(os.path.join(REF, "speech", "models", "seq2seq.py")).read()
    broken = "beam = filter(lambda x : x[0][-1] != end_tok, new_beam)"
    assert broken in src
    src = src.replace(broken, "beam = list(filter(lambda x : x[0][-1] != end_tok, new_beam))")
    m = types.ModuleType("refmodels.seq2seq")
    m.__package__ = "refmodels"
    sys.modules["refmodels.seq2seq"] = m
    exec(compile(src, "seq2seq.py(py3-fixed in memory)", "exec"), m.__dict__)
    mods["seq2seq"] = m
    return mods


def sd_to_np(sd, prefix):
    return {prefix + k.replace(".", "__"): v.detach().numpy() for k, v in sd.items()}


def main():
    mods = load_ref_models()
    dec = mods["ctc_decoder"].decode

    # ---------------- ctc_decode ----------------
    g = {}
    np.random.seed(3)
    probs = np.random.rand(50, 20)
    probs = probs / np.sum(probs, axis=1, keepdims=True)
    g["demo_probs"] = probs
    for beam in (10, 8, 1):
        lab, sc = dec(probs, beam_size=beam, blank=0)
        g["demo_b%d_labels" % beam] = np.array(lab, np.int32)
        g["demo_b%d_score" % beam] = np.float64(sc)
    rng = np.random.RandomState(11)
    # float32 probabilities, blank = last class (what CTC.infer feeds, ctc_model.py:55-60)
    x = rng.randn(60, 12).astype(np.float32) * 2.0
    p32 = np.exp(x - x.max(1, keepdims=True))
    p32 = (p32 / p32.sum(1, keepdims=True)).astype(np.float32)
    g["f32_probs"] = p32
    for beam in (1, 4, 10):
        lab, sc = dec(p32.astype(np.float64), beam_size=beam, blank=11)
        g["f32_b%d_labels" % beam] = np.array(lab, np.int32)
        g["f32_b%d_score" % beam] = np.float64(sc)
    # peaky distribution with many repeats (exercises the merge branch, ctc_decoder.py:87-103)
    T, S = 40, 5
    pk = np.full((T, S), 0.02)
    seq = rng.randint(0, S, size=T)
    seq[5:12] = 2
    seq[20:26] = 3
    pk[np.arange(T), seq] = 0.92
    pk = pk / pk.sum(1, keepdims=True)
    g["peaky_probs"] = pk
    for beam in (1, 3, 8):
        lab, sc = dec(pk, beam_size=beam, blank=4)
        g["peaky_b%d_labels" % beam] = np.array(lab, np.int32)
        g["peaky_b%d_score" % beam] = np.float64(sc)
    # exact ties: uniform distribution -> every candidate has the same score each step
    un = np.full((6, 4), 0.25)
    g["tie_probs"] = un
    for beam in (1, 3):
        lab, sc = dec(un, beam_size=beam, blank=0)
        g["tie_b%d_labels" % beam] = np.array(lab, np.int32)
        g["tie_b%d_score" % beam] = np.float64(sc)
    np.savez(os.path.join(OUT, "ctc_decode.npz"), **g)

    # ---------------- encoder ----------------
    Model = mods["model"].Model
    g = {}
    tiny = {"dropout": 0.0, "encoder": {"conv": [[32, 5, 32, 2]],
                                        "rnn": {"dim": 16, "bidirectional": False, "layers": 1}}}
    bi = {"dropout": 0.0, "encoder": {"conv": [[8, 5, 8, 2], [8, 5, 8, 2]],
                                      "rnn": {"dim": 32, "bidirectional": True, "layers": 2}}}
    for tag, cfg, fdim, T in (("tiny", tiny, 40, 100), ("bi", bi, 80, 61)):
        torch.manual_seed(0)
        m = Model(fdim, cfg)
        x = torch.randn(4, T, fdim)
        with torch.no_grad():
            y = m.encode(x)
        g.update(sd_to_np(m.state_dict(), tag + "__sd__"))
        g[tag + "_x"] = x.numpy()
        g[tag + "_y"] = y.numpy()
    np.savez(os.path.join(OUT, "encoder_tiny.npz"), **g)

    # ---------------- CTC model logits (encoder + fc; loss is external) ----------------
    # ctc_model.py imports functions.ctc at module level; restate only what forward_impl does
    # with the reference's own Model + LinearND (ctc_model.py:19,25-32).
    torch.manual_seed(0)
    m = Model(40, tiny)
    fc = mods["model"].LinearND(16, 11)
    x = torch.randn(4, 100, 40)
    with torch.no_grad():
        logits = fc(m.encode(x))
    g = sd_to_np(m.state_dict(), "sd__")
    g.update(sd_to_np(fc.state_dict(), "sd__fc__"))
    g["x"] = x.numpy()
    g["logits"] = logits.numpy()
    np.savez(os.path.join(OUT, "ctc_tiny.npz"), **g)

    # ---------------- Seq2Seq ----------------
    Seq2Seq = mods["seq2seq"].Seq2Seq
    np.random.seed(1337)
    torch.manual_seed(1337)
    conf = {"dropout": 0.0, "encoder": {"conv": [[32, 5, 32, 2]],
                                        "rnn": {"dim": 16, "bidirectional": False, "layers": 1}},
            "decoder": {"embedding_dim": 16, "layers": 2}}
    vocab = 10
    s2s = Seq2Seq(120, vocab + 1, conf)
    inputs = [np.random.randn(100, 120) for _ in range(4)]
    # labels: <s>=10 (last index) first, </s>=9 last, like Preprocessor.encode with start_and_end
    labels = [[10] + list(np.random.randint(0, 9, 18)) + [9] for _ in range(4)]
    batch = (inputs, labels)
    with torch.no_grad():
        out = s2s(batch)
        loss = s2s.loss(batch)
        x, y = s2s.collate(*batch)
        x_enc = s2s.encode(x)
        out2, aligns = s2s.decode(x_enc, y)
        greedy = s2s.infer(batch, max_len=30)
        beams = {}
        for e in range(4):
            one = ([inputs[e]], [labels[e]])
            for bs in (1, 3, 8):
                beams[(e, bs)] = s2s.beam_search(one, beam_size=bs, max_len=30)[0]
    g = sd_to_np(s2s.state_dict(), "sd__")
    g["inputs"] = np.stack(inputs).astype(np.float32)
    g["labels"] = np.array(labels, np.int64)
    g["out"] = out.numpy()
    g["loss"] = np.float64(loss.item())
    g["x_enc"] = x_enc.numpy()
    g["aligns"] = aligns.numpy()
    g["greedy"] = np.array([r + [-1] * (40 - len(r)) for r in greedy], np.int64)
    for (e, bs), hyp in beams.items():
        g["beam_u%d_b%d" % (e, bs)] = np.array(hyp, np.int64)
    np.savez(os.path.join(OUT, "seq2seq_tiny.npz"), **g)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print("  %-20s %7d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
