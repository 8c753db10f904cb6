"""Developer tool: forward GRU step time under timing ablations (results are wrong by design)."""
import sys, torch
sys.path.insert(0, ".")
from speech_b200 import _lib, ops
lib = _lib.load()
torch.manual_seed(0)
B, T, In, H = 64, 247, 2048, 1024
rnn = torch.nn.GRU(In, H, 1, batch_first=True, bidirectional=True).cuda()
x = torch.randn(B, T, In, device="cuda")
names = {0: "baseline", 1: "no proxy fence", 2: "no off-path stores", 3: "no fence + no off-path",
         4: "no TMA loads", 8: "no MMA (no accfull wait)", 12: "no TMA, no MMA",
         16: "no grid-barrier wait", 28: "no barrier wait, no TMA, no MMA",
         31: "epilogue + arrive only"}
sel = [int(a) for a in sys.argv[1:]] or list(names)
for flags, name in names.items():
    if flags not in sel:
        continue
    lib.sb_debug_gru_flags(flags)
    with torch.no_grad():
        for _ in range(2):
            ops.gru_stack(x, rnn)
        ops.profile_begin()
        for _ in range(3):
            ops.gru_stack(x, rnn)
        prof = ops.profile_end()
    n, ms, _ = prof["gru_fwd"]
    print("%-36s %.2f us/step" % (name, ms / n / T * 1e3), flush=True)
lib.sb_debug_gru_flags(0)
