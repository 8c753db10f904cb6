"""Developer tool: run every kernel once at small sizes (for compute-sanitizer).
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from speech_b200.models import CTC, Seq2Seq, Transducer
from speech_b200.optim import FlatSGD
torch.manual_seed(0); np.random.seed(0)
cfg = {"dropout": 0.0, "encoder": {"conv": [[8, 5, 8, 2], [8, 5, 8, 2]],
                                   "rnn": {"dim": 256, "bidirectional": True, "layers": 2}},
       "decoder": {"embedding_dim": 256, "layers": 1}}
inputs = [np.random.randn(90 + 3 * i, 80).astype(np.float32) for i in range(5)]
labels = [np.random.randint(0, 28, 7 + i).tolist() for i in range(5)]
m = CTC(80, 28, cfg).cuda()
opt = FlatSGD(m, lr=1e-3)
for _ in range(2):
    opt.zero_grad(); loss = m.loss((inputs, labels)); loss.backward(); opt.step()
print("ctc step ok", loss.item())
print("infer", [len(p) for p in m.infer((inputs, labels), beam_size=4)])
s = Seq2Seq(80, 30, cfg).cuda()
lab2 = [[29] + l + [28] for l in labels]
l2 = s.loss((inputs, lab2)); l2.backward()
print("seq2seq ok", l2.item(), [len(h) for h in s.infer((inputs, lab2), max_len=12)],
      s.beam_search(([inputs[0]], [lab2[0]]), beam_size=3, max_len=8))
t = Transducer(80, 28, cfg).cuda()
l3 = t.loss((inputs, labels)); l3.backward()
print("transducer ok", l3.item())
torch.cuda.synchronize()
