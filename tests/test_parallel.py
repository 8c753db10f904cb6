"""CPU (gloo, world_size 2) test of the data-parallel gradient synchronisation:
the N-rank sharded run must reproduce the 1-rank gradients of the global batch (sum reduction)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from speech_b200.parallel import GradSync
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    sync = GradSync(model, world)
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    items = list(range(8))
    mine = sync.shard(items, rank)
    model.zero_grad(set_to_none=False)
    loss = model(x[mine]).pow(2).sum()      # SUM over the shard, like the reference's CTC loss
    loss.backward()
    # gradients accumulate in place into the flat buffer views
    assert all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in model.parameters())
    sync.all_reduce()
    ret[rank] = sync.flat.clone()
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_rank_global_batch():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    model(x).pow(2).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(ret[0], ref, rtol=1e-5, atol=1e-6)
    assert torch.equal(ret[0], ret[1])


def _bucket_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from speech_b200.parallel import GradSync
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4),
                                torch.nn.Tanh(), torch.nn.Linear(4, 3))
    sync = GradSync(model, world)
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    mine = sync.shard(list(range(8)), rank)
    model.zero_grad(set_to_none=False)
    model(x[mine]).pow(2).sum().backward()
    # announce the layers out of order and only some of them, as the backward pass would:
    # last layer first, then the middle one; the first layer is left for finish() to pick up
    sync.ready(list(model[4].parameters()))
    sync.ready(list(model[2].parameters()))
    sync.all_reduce()
    ret[rank] = sync.flat.clone()
    # a second step must start from a clean slate (no stale pending work)
    model.zero_grad(set_to_none=False)
    model(x[mine]).pow(2).sum().backward()
    sync.all_reduce()
    ret[rank + world] = sync.flat.clone()
    dist.destroy_process_group()


def test_bucketed_overlapped_all_reduce_equals_single_call():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, port, ret), nprocs=world, join=True)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4),
                                torch.nn.Tanh(), torch.nn.Linear(4, 3))
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    model(x).pow(2).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    for k in range(4):
        assert torch.allclose(ret[k], ref, rtol=1e-5, atol=1e-6), k
    assert torch.equal(ret[0], ret[1])


def test_bucket_reducer_rejects_overlapping_announcements():
    from speech_b200.parallel import BucketReducer
    import pytest

    class _W:
        def wait(self):
            pass
    r = BucketReducer(torch.zeros(16), world_size=2)
    r.pending = [(0, 8, _W()), (4, 12, _W())]
    with pytest.raises(RuntimeError):
        r.finish()
