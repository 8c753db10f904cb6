"""GPU parity of ALL parameter gradients at north-star depth (5-layer biGRU-1024, T=1000 -> T'=247,
WSJ conv stack, CTC loss; B=8 utterances of the bench batch) against the reference model restated
on the CPU in fp32 (oracle/model_ref.py = the reference's nn.Conv2d / nn.GRU / nn.Linear +
log_softmax + F.ctc_loss, speech/models/model.py:60-79, ctc_model.py:25-40).

The north star asks for "CTC loss and gradients"; the kernels round GEMM / recurrent operands to
bf16 (fp32 accumulate), so the comparison reports, per parameter, the relative L2 error and the
cosine to the fp32 reference through 247 x 5 recurrent steps.  A table is written to
gpurun_out/northstar_grads.txt when that directory exists."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _northstar(nutt=8):
    sys.path.insert(0, ROOT)
    import bench
    from oracle.model_ref import RefCTC
    from speech_b200.models import CTC
    torch.manual_seed(0)
    m = CTC(bench.F_IN, bench.VOCAB, bench.MODEL_CFG)
    ref = RefCTC(bench.F_IN, bench.VOCAB, bench.MODEL_CFG)
    ref.load_from_dropin({k: v.detach().clone() for k, v in m.state_dict().items()})
    inputs, labels = bench.synth_batch(nutt)
    return m, ref, inputs, labels


def _ref_grads(ref, inputs, labels):
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    x = torch.from_numpy(np.stack(inputs))
    flat = torch.tensor([t for l in labels for t in l], dtype=torch.int32)
    lens = torch.tensor([len(l) for l in labels], dtype=torch.int32)
    loss = ref.loss(x, flat, lens)
    loss.backward()
    out = {}
    for k, p in ref.named_parameters():
        name = k[len("enc."):] if k.startswith("enc.") else "fc.fc." + k[len("fc."):]
        out[name] = p.grad.detach().double()
    return float(loss.item()), out


def _report(rows, tag):
    lines = ["%-28s %10s %10s %12s" % ("parameter (" + tag + ")", "rel L2", "1-cos", "|ref| L2")]
    for n, rel, cos, nrm in rows:
        lines.append("%-28s %10.3e %10.3e %12.4e" % (n, rel, 1.0 - cos, nrm))
    txt = "\n".join(lines)
    print(txt)
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "northstar_grads_%s.txt" % tag), "w") as fh:
            fh.write(txt + "\n")


def _compare(m, ref_grads):
    rows = []
    for n, p in m.named_parameters():
        g = p.grad.detach().double().cpu()
        r = ref_grads[n]
        rel = ((g - r).norm() / r.norm()).item()
        cos = (torch.dot(g.flatten(), r.flatten()) / (g.norm() * r.norm())).item()
        rows.append((n, rel, cos, r.norm().item()))
    return rows


def test_all_parameter_gradients_at_northstar_depth(cuda_lib):
    m, ref, inputs, labels = _northstar()
    ref_loss, ref_grads = _ref_grads(ref, inputs, labels)
    m.cuda()
    m.set_train()
    loss = m.loss((tuple(inputs), tuple(labels)))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss) / ref_loss < 1e-4, (loss.item(), ref_loss)
    rows = _compare(m, ref_grads)
    _report(rows, "bf16")
    # bf16 operands, fp32 accumulate, through 5 layers x 247 steps: every parameter's gradient
    # keeps its direction (cosine) and its size (relative L2) against the fp32 reference
    worst_rel = max(r[1] for r in rows)
    worst_cos = min(r[2] for r in rows)
    assert worst_cos > 0.999, rows
    assert worst_rel < 5e-2, rows


def test_parity_mode_loss_matches_fp32_reference_at_northstar_depth(cuda_lib):
    """`model.parity_mode = True` (inference only): split-bf16 3-pass GEMMs + the fp32 recurrence
    kernel (csrc/gru_f32.cu).  The CTC loss of the north-star model must then agree with the fp32
    CPU reference to 1e-5 relative and the logits to 5e-4 absolute - the bf16 operand path is
    held to 1e-4 / ~1e-2; this is the measuring stick SURVEY.md section 7 asks for."""
    import bench
    m, ref, inputs, labels = _northstar(4)
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    x = torch.from_numpy(np.stack(inputs))
    flat = torch.tensor([t for l in labels for t in l], dtype=torch.int32)
    lens = torch.tensor([len(l) for l in labels], dtype=torch.int32)
    with torch.no_grad():
        ref_logits = ref.logits(x)
        ref_loss = float(ref.loss(x, flat, lens).item())
    m.cuda()
    m.set_eval()
    batch = (tuple(inputs), tuple(labels))
    with torch.no_grad():
        bf16_logits = m(batch)
        bf16_loss = float(m.loss(batch).item())
        m.parity_mode = True
        par_logits = m(batch)
        par_loss = float(m.loss(batch).item())
    e_bf16 = (bf16_logits.cpu() - ref_logits).abs().max().item()
    e_par = (par_logits.cpu() - ref_logits).abs().max().item()
    d = os.path.join(ROOT, "gpurun_out")
    msg = ("north-star model, 4 utterances: |logits - fp32 ref| max: bf16 path %.3e, parity mode %.3e; "
           "CTC loss rel delta: bf16 path %.3e, parity mode %.3e"
           % (e_bf16, e_par, abs(bf16_loss - ref_loss) / ref_loss, abs(par_loss - ref_loss) / ref_loss))
    print(msg)
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_mode.txt"), "w") as fh:
            fh.write(msg + "\n")
    assert e_par < 5e-4, msg
    assert abs(par_loss - ref_loss) / ref_loss < 1e-5, msg
    assert abs(bf16_loss - ref_loss) / ref_loss < 1e-4, msg
