"""Developer tool: where one training step of the WSJ attention model / the RNN-T goes (per-kernel device time
from bracketing CUDA events, host enqueue time of the step, step time without the profiler)."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from speech_b200 import ops
from speech_b200.models import Seq2Seq
from speech_b200.optim import FlatSGD

which = sys.argv[1] if len(sys.argv) > 1 else "wsj"
torch.manual_seed(0)
rng = np.random.RandomState(2)
if which == "wsj":
    cfg = {"dropout": 0.0, "encoder": {"conv": bench.WSJ_CONV,
                                       "rnn": {"dim": 512, "bidirectional": True, "layers": 3}},
           "decoder": {"embedding_dim": 512, "layers": 1, "log_t": True}}
    V = 32
    m = Seq2Seq(bench.F_IN, V, cfg).cuda()
    lab = lambda: [V - 1] + rng.randint(0, V - 2, size=rng.randint(60, 100)).tolist() + [V - 2]
    batch = (tuple(rng.randn(800, bench.F_IN).astype(np.float32) for _ in range(16)),
             tuple(lab() for _ in range(16)))
elif which == "rnnt":
    m, batch = bench.rnnt_workload(32, 0)
else:
    raise SystemExit("usage: step_breakdown.py [wsj|rnnt]")
print("config:", which)
m.set_train()
opt = FlatSGD(m, lr=1e-4, momentum=0.0, max_grad_norm=200.0)


def step():
    opt.zero_grad()
    loss = m.loss(batch)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms, step wall %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
ops.profile_begin()
step()
torch.cuda.synchronize()
rec = ops.profile_end()
tot = sum(v[1] for v in rec.values())
print("sum of bracketed launches: %.2f ms" % tot)
for k, (n, ms, _) in sorted(rec.items(), key=lambda kv: -kv[1][1]):
    print("  %-24s n=%4d  %8.3f ms  (%.1f us each)" % (k, n, ms, 1e3 * ms / n))
