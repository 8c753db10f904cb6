import sys, torch
sys.path.insert(0, "/root/repo")
from speech_b200 import ops, _lib
from speech_b200.ops import gru_stack
lib = _lib.load()
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(3)
rnn = torch.nn.GRU(64, H, 2, batch_first=True, bidirectional=True)
x = torch.randn(8, 7, 64)
r64 = torch.nn.GRU(64, H, 2, batch_first=True, bidirectional=True).double()
r64.load_state_dict({k: v.double() for k, v in rnn.state_dict().items()})
y64, _ = r64(x.double()); y64.sum().backward()
ref = {n: p.grad.float() for n, p in r64.named_parameters()}
rnn = rnn.cuda(); xc = x.cuda()
for it in range(5):
    for p in rnn.parameters(): p.grad = None
    y = gru_stack(xc, rnn); y.sum().backward(); torch.cuda.synchronize()
    worst = max(((p.grad.cpu() - ref[n]).abs().max() / ref[n].abs().max()).item() for n, p in rnn.named_parameters())
    print("H=%d call %d: worst grad rel err vs fp64 reference %.3e" % (H, it, worst))
