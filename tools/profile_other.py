"""Developer tool for ncu: one training step + one decode of the attention and transducer models
(the kernels the north-star bench step does not launch): joint_kernel, rnnt_fwd_bwd_kernel,
rnnt_decode_static_kernel, s2s_cell_fwd/bwd, s2s_attn_fwd/bwd, s2s_beam_select."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from speech_b200.models import Seq2Seq

m, batch = bench.rnnt_workload(8, 0)
for _ in range(2):
    m.zero_grad()
    m.loss(batch).backward()
m.set_eval()
m.infer(batch, beam_size=4)
torch.cuda.synchronize()
torch.manual_seed(0)
cfg = {"dropout": 0.0, "encoder": {"conv": bench.WSJ_CONV,
                                   "rnn": {"dim": 512, "bidirectional": True, "layers": 3}},
       "decoder": {"embedding_dim": 512, "layers": 1, "log_t": True}}
V = 32
s2s = Seq2Seq(bench.F_IN, V, cfg).cuda()
rng = np.random.RandomState(0)
lab = lambda: [V - 1] + rng.randint(0, V - 2, size=30).tolist() + [V - 2]
b2 = (tuple(rng.randn(800, bench.F_IN).astype(np.float32) for _ in range(16)), tuple(lab() for _ in range(16)))
for _ in range(2):
    s2s.zero_grad()
    s2s.loss(b2).backward()
s2s.set_eval()
s2s.infer(b2, max_len=20)
s2s.beam_search(((b2[0][0],), (b2[1][0],)), beam_size=8, max_len=20)
torch.cuda.synchronize()
print("ok")
