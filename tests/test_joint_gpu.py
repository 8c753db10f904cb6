"""GPU parity of the fused RNN-T joint + loss (csrc/joint.cu + the compact lattice mode of
csrc/rnnt.cu) against the reference's formulation restated on the CPU in float64:

    lp = log_softmax(fc2(relu(fx[:, :, None] + fy[:, None])), 3)      transducer_model.py:71-76
    loss = TransducerLoss(lp, labels, x_lens, y_lens)                  (oracle/rnnt_ref.py)

The kernels round the hidden activations and fc2's weight to bf16 (fp32 accumulate); the reference
here applies the SAME operand rounding so that the comparison isolates the kernels:
loss within 1e-4 relative, every gradient within 2 % of its largest entry (bf16 rounding of the
backward operands), and the full log-probability tensor `infer` uses within 1e-4 absolute."""
import numpy as np
import pytest
import torch

from oracle import rnnt_ref

pytestmark = pytest.mark.gpu


def _rnd(t):     # bf16 rounding with a straight-through gradient
    return t + (t.detach().float().bfloat16().double() - t.detach())


def _case(B, T, U1, H, V1, seed):
    rng = np.random.RandomState(seed)
    fx = torch.from_numpy(rng.randn(B, T, H) * 0.7)
    fy = torch.from_numpy(rng.randn(B, U1, H) * 0.7)
    w2 = torch.from_numpy(rng.randn(V1, H) / np.sqrt(H))
    b2 = torch.from_numpy(rng.randn(V1) * 0.1)
    ylen = rng.randint(max(0, U1 - 4), U1, size=B).astype(np.int32)
    ylen[0] = U1 - 1
    xlen = np.full(B, T, np.int32)
    ymat = rng.randint(0, V1 - 1, size=(B, U1 - 1)).astype(np.int64)
    flat = np.concatenate([ymat[b, :ylen[b]] for b in range(B)]).astype(np.int32)
    return fx, fy, w2, b2, ymat, flat, xlen, ylen


@pytest.mark.parametrize("B,T,U1,H,V1,seed", [
    (2, 5, 4, 16, 7, 0),          # tiny: H < one 64-column block
    (3, 23, 9, 96, 11, 1),        # ragged labels, H not a multiple of 64, several slabs of frames
    (4, 40, 13, 256, 29, 2),      # shipped width class, V+1 = 29
    (2, 17, 6, 128, 62, 3),       # TIMIT-sized vocabulary (V+1 = 62 -> 64-column accumulator)
])
def test_fused_joint_loss_matches_reference_formulation(cuda_lib, B, T, U1, H, V1, seed):
    from speech_b200.functions.transducer import JointTransducerLoss, joint_log_probs
    fx, fy, w2, b2, ymat, flat, xlen, ylen = _case(B, T, U1, H, V1, seed)
    blank = V1 - 1
    # ---- reference (fp64, same operand rounding) ----
    rf = [t.clone().requires_grad_(True) for t in (fx, fy, w2, b2)]
    z = _rnd(torch.relu(rf[0][:, :, None, :] + rf[1][:, None, :, :]))
    lp = torch.log_softmax(z @ _rnd(rf[2]).t() + rf[3], 3)
    costs, g = rnnt_ref.rnnt_loss_and_grad(lp.detach().numpy(), flat, xlen, ylen, blank)
    lp.backward(torch.from_numpy(g))
    # ---- ours ----
    mine = [t.float().cuda().requires_grad_(True) for t in (fx, fy, w2, b2)]
    loss = JointTransducerLoss(blank=blank)(mine[0], mine[1], mine[2], mine[3],
                                            torch.from_numpy(ymat), torch.from_numpy(flat),
                                            torch.from_numpy(xlen), torch.from_numpy(ylen))
    assert loss.shape == (1,)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - costs.sum()) / costs.sum() < 1e-4, (loss.item(), costs.sum())
    for name, a, r in zip(("fx", "fy", "w2", "b2"), mine, rf):
        ref = r.grad
        err = (a.grad.double().cpu() - ref).abs().max().item()
        assert err < 2e-2 * ref.abs().max().item() + 1e-6, (name, err, ref.abs().max().item())
    # ---- the full log-probability tensor of `forward` / `infer` ----
    class FC2:
        weight, bias = mine[2].detach(), mine[3].detach()
    full = joint_log_probs(mine[0].detach(), mine[1].detach(), FC2, torch.from_numpy(ymat), blank)
    assert tuple(full.shape) == (B, T, U1, V1)
    assert (full.double().cpu() - lp.detach()).abs().max().item() < 2e-3
    assert np.allclose(np.exp(full.double().cpu().numpy()).sum(-1), 1.0, atol=1e-5)


def test_transducer_model_trains_through_the_fused_joint(cuda_lib):
    """Transducer.loss (drop-in class): finite loss, every parameter receives a gradient, and the
    loss equals the reference formulation evaluated on the model's own full log-probabilities."""
    from speech_b200.models import Transducer
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = {"dropout": 0.0, "encoder": {"conv": [[8, 5, 8, 2]],
                                       "rnn": {"dim": 32, "bidirectional": True, "layers": 1}},
           "decoder": {"embedding_dim": 16, "layers": 1}}
    m = Transducer(40, 10, cfg).cuda()
    inputs = [np.random.randn(60, 40).astype(np.float32) for _ in range(3)]
    labels = [np.random.randint(0, 10, n).tolist() for n in (5, 3, 4)]
    batch = (inputs, labels)
    loss = m.loss(batch)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all()
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        assert p.grad.abs().sum() > 0, n
    with torch.no_grad():
        lp = m(batch)                                     # (B, T', U+1, V+1)
    x, y, x_lens, y_lens = m.collate(*batch)
    costs, _ = rnnt_ref.rnnt_loss_and_grad(lp.double().cpu().numpy(), y.numpy(), x_lens.numpy(),
                                           y_lens.numpy(), m.blank)
    assert abs(loss.item() - costs.sum()) / costs.sum() < 1e-4
    preds = m.infer(batch)
    assert len(preds) == 3 and all(isinstance(p, list) for p in preds)
