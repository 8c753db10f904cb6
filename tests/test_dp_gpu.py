"""GPU (NCCL, 2 ranks) test of the data-parallel step: the minibatch is sharded, every GRU layer's
gradients are all-reduced from inside the backward pass (overlapped buckets), and the result must
equal the single-GPU gradients of the whole minibatch.  Skipped on a box with one GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CFG = {"dropout": 0,
       "encoder": {"conv": [[8, 5, 8, 2]],
                   "rnn": {"dim": 64, "bidirectional": True, "layers": 3}}}


def _batch():
    rng = np.random.RandomState(0)
    inputs = [rng.randn(60, 40).astype(np.float32) for _ in range(8)]
    labels = [rng.randint(0, 10, size=rng.randint(3, 9)).tolist() for _ in range(8)]
    return inputs, labels


def _model():
    from speech_b200 import models
    torch.manual_seed(0)
    return models.CTC(40, 10, CFG).cuda()


def _worker(rank, world, port, overlap, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from speech_b200 import ops
    from speech_b200.optim import FlatSGD
    model = _model()
    opt = FlatSGD(model, lr=1e-3, max_grad_norm=1e9, world_size=world, overlap=overlap)
    inputs, labels = _batch()
    per = len(inputs) // world
    sl = slice(rank * per, (rank + 1) * per)
    for _ in range(2):                       # twice: no stale buckets across steps
        opt.zero_grad()
        loss = model.loss((inputs[sl], labels[sl]))
        loss.backward()
        n_pending = len(opt.reducer.pending)
        opt.reducer.finish()
    torch.cuda.synchronize()
    ret[rank] = (opt.flat_g.cpu(), n_pending)
    ops.set_grad_ready_hook(None)
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_gpu_overlapped_all_reduce_matches_single_gpu(cuda_lib, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from speech_b200.optim import FlatSGD
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, overlap, ret), nprocs=2, join=True)
    model = _model()
    opt = FlatSGD(model, lr=1e-3, max_grad_norm=1e9)
    opt.zero_grad()
    model.loss(_batch()).backward()
    ref = opt.flat_g.cpu()
    g0, pending0 = ret[0]
    g1, _ = ret[1]
    assert torch.equal(g0, g1)
    assert pending0 == (3 if overlap else 0)          # one bucket per GRU layer
    # same per-utterance arithmetic, different fp32 summation order over the minibatch
    assert (g0 - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
