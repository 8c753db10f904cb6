"""Debug helper: time each stage of one north-star training step with syncs (not a benchmark)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from speech_b200 import ops, _lib
from speech_b200.models import CTC
import bench

def tick(msg, t0):
    torch.cuda.synchronize()
    print("[%.3fs] %s" % (time.perf_counter() - t0, msg), flush=True)

L = int(sys.argv[1]) if len(sys.argv) > 1 else 5
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfg = {"dropout": 0.0, "encoder": {"conv": [[32, 5, 8, 2], [32, 5, 8, 2]],
                                   "rnn": {"dim": 1024, "bidirectional": True, "layers": L}}}
t0 = time.perf_counter()
torch.manual_seed(0)
m = CTC(80, 28, cfg).cuda()
tick("model on gpu", t0)
rng = np.random.RandomState(0)
inputs = [rng.randn(T, 80).astype(np.float32) for _ in range(64)]
labels = [rng.randint(0, 28, size=rng.randint(10, 31)).tolist() for _ in range(64)]
x, y, xl, yl = m.collate(inputs, labels)
x = x.cuda()
tick("batch on gpu", t0)
for it in range(3):
    ops._prof_detail = (it == 2)
    ops.profile_begin()
    c = ops.conv_stack(x, m.conv, True)
    tick("conv", t0)
    h = ops.gru_stack(c, m.rnn)
    tick("gru fwd", t0)
    out = m.fc(h[:, :, :1024] + h[:, :, 1024:])
    loss = m.ctc_loss(out, y, xl, yl)
    tick("fc+ctc loss=%.3f" % loss.item(), t0)
    loss.backward()
    tick("backward", t0)
    for k, v in sorted(ops.profile_end().items(), key=lambda kv: -kv[1][1]):
        print("    %-34s n=%3d  %.3f ms  %.1f TFLOP/s" % (k, v[0], v[1], v[2] / 1e9 / max(v[1], 1e-9)))
