"""Data-parallel gradient synchronisation: one process per GPU, minibatch sharded B/N per rank,
ONE summed all-reduce of the gradients per step over NCCL (NVLink 5 / NVSwitch).

The reference has no multi-GPU path at all (train.py:92); this adds exactly the
collective the north star names.  Gradients are SUMMED (not averaged): the reference's loss is a
sum over the minibatch (ctc_model.py:38-39), so the sum over ranks of per-shard gradients equals
the single-GPU gradient of the global batch.

All parameter gradients are views into one flat fp32 buffer, so the all-reduce is a single NCCL
call (339.5 MB at the north-star config) and `zero_grad(set_to_none=False)` is one memset.
"""
import torch


class GradSync:
    def __init__(self, model, world_size, backend_group=None):
        self.world = world_size
        self.group = backend_group
        self.params = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def all_reduce(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)

    def shard(self, items, rank):
        """contiguous B/N shard of a per-utterance list for `rank`."""
        per = len(items) // self.world
        return items[rank * per:(rank + 1) * per]
