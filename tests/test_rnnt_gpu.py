"""GPU parity: sb_rnnt_fwd_bwd (through the C ABI) vs the float64 CPU oracle.
Bar: loss and gradients within 1e-4 relative (gradient relative to its largest entry)."""
import numpy as np
import pytest
import torch

from oracle import rnnt_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T,U,V,seed", [(2, 3, 2, 4, 0), (4, 20, 7, 11, 1), (3, 48, 20, 11, 2),
                                          (2, 60, 270, 6, 3)])
def test_rnnt_matches_oracle(cuda_lib, B, T, U, V, seed):
    from speech_b200.functions.transducer import rnnt_costs_and_grads
    rng = np.random.RandomState(seed)
    x = rng.randn(B, T, U + 1, V).astype(np.float32)
    lp = torch.log_softmax(torch.from_numpy(x), 3)
    ylen = rng.randint(max(0, U - 5), U + 1, size=B).astype(np.int32)
    ylen[0] = U
    xlen = rng.randint(max(1, T - 4), T + 1, size=B).astype(np.int32)
    xlen[0] = T
    flat = np.concatenate([rng.randint(0, V - 1, size=n) for n in ylen]).astype(np.int32)
    c, g = rnnt_costs_and_grads(lp.cuda(), torch.from_numpy(flat), torch.from_numpy(xlen),
                                torch.from_numpy(ylen))
    c_ref, g_ref = rnnt_ref.rnnt_loss_and_grad(lp.numpy(), flat, xlen, ylen)
    np.testing.assert_allclose(c.cpu().numpy(), c_ref, rtol=1e-4)
    g = g.cpu().numpy()
    for b in range(B):
        assert np.abs(g[b] - g_ref[b]).max() / max(np.abs(g_ref[b]).max(), 1e-9) < 1e-4


def test_transducer_loss_module_contract(cuda_lib):
    from speech_b200.functions.transducer import TransducerLoss
    rng = np.random.RandomState(5)
    x = torch.from_numpy(rng.randn(2, 6, 4, 5).astype(np.float32)).cuda().requires_grad_(True)
    lp = torch.log_softmax(x, 3)
    flat = torch.IntTensor([1, 2, 3, 0, 1])
    loss = TransducerLoss()(lp, flat, torch.IntTensor([6, 6]), torch.IntTensor([3, 2]))
    assert loss.shape == (1,)
    loss.backward()
    # d loss / d logits through log_softmax: every (t,u) row of the gradient sums to ~0
    assert x.grad.abs().sum() > 0
    c_ref, _ = rnnt_ref.rnnt_loss_and_grad(lp.detach().cpu().numpy(), flat.numpy(), [6, 6], [3, 2])
    assert abs(loss.item() - c_ref.sum()) / c_ref.sum() < 1e-4
