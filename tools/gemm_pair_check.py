"""Developer check: CTA-pair (cta_group::2) GEMM against the single-CTA kernel - values and time."""
import sys
import torch
from speech_b200 import _lib, ops

lib = _lib.load()
torch.manual_seed(0)
dev = "cuda"


def run(A, B, mode, bias=None, out=None, accumulate=False, split_k=1, iters=0):
    lib.sb_debug_gemm_mt1(1 | (4 if mode else 0))
    C = ops.gemm_bf16_tn(A, B, out=out, bias=bias, accumulate=accumulate, split_k=split_k)
    if iters:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.gemm_bf16_tn(A, B, out=out, bias=bias, accumulate=accumulate, split_k=split_k)
        e1.record()
        torch.cuda.synchronize()
        return C, e0.elapsed_time(e1) / iters
    torch.cuda.synchronize()
    return C, None


ok = True
for (M, N, K, use_bias) in [(512, 512, 256, False), (16000, 6144, 2048, True), (1000, 700, 192, True),
                            (16000, 2048, 6144, False), (15808, 6144, 480, True)]:
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev) if use_bias else None
    C1, _ = run(A, B, 0, bias)
    C2, _ = run(A, B, 1, bias)
    err = (C1 - C2).abs().max().item()
    print("M=%d N=%d K=%d bias=%s  max|pair-single| = %.3e" % (M, N, K, use_bias, err), flush=True)
    ok = ok and err < 1e-3
# accumulate + split-K (weight-gradient shape)
M, N, K = 6144, 2048, 16000
A = torch.randn(M, K, device=dev).bfloat16()
B = torch.randn(N, K, device=dev).bfloat16()
o1 = torch.ones(M, N, device=dev)
o2 = torch.ones(M, N, device=dev)
run(A, B, 0, out=o1, accumulate=True, split_k=4)
run(A, B, 1, out=o2, accumulate=True, split_k=4)
err = (o1 - o2).abs().max().item() / o1.abs().max().item()
print("wgrad split-k accumulate rel err %.3e" % err, flush=True)
ok = ok and err < 1e-4
if not ok:
    print("MISMATCH")
    sys.exit(1)
for (M, N, K) in [(16000, 6144, 2048), (16000, 2048, 6144), (16000, 6144, 480), (8000, 6144, 2048)]:
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    for mode in (0, 1):
        _, ms = run(A, B, mode, iters=20)
        print("M=%d N=%d K=%d %s: %.3f ms  %.0f TF/s" % (M, N, K, "pair  " if mode else "single",
                                                          ms, 2.0 * M * N * K / ms / 1e9), flush=True)
for (M, N, K, sk) in [(6144, 2048, 16000, 2), (3072, 2048, 16000, 2), (2048, 1024, 16000, 4),
                      (1024, 1024, 16000, 9), (3072, 480, 16000, 1)]:
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    o1 = torch.zeros(M, N, device=dev)
    o2 = torch.zeros(M, N, device=dev)
    run(A, B, 0, out=o1, accumulate=True, split_k=sk)
    run(A, B, 1, out=o2, accumulate=True, split_k=sk)
    print("wgrad M=%d N=%d rel err %.2e" % (M, N, (o1 - o2).abs().max().item() / o1.abs().max().item()))
    for mode in (0, 1):
        _, ms = run(A, B, mode, out=o1, accumulate=True, split_k=sk, iters=20)
        print("wgrad M=%d N=%d K=%d split%d %s: %.3f ms  %.0f TF/s" % (
            M, N, K, sk, "pair  " if mode else "single", ms, 2.0 * M * N * K / ms / 1e9), flush=True)

# is the accumulate (reduce-add) epilogue what holds the weight-gradient GEMMs back?
M, N, K = 6144, 2048, 16000
A = torch.randn(M, K, device=dev).bfloat16()
B = torch.randn(N, K, device=dev).bfloat16()
o = torch.zeros(M, N, device=dev)
for mode in (0, 1):
    for acc, sk in ((False, 1), (True, 1), (True, 2)):
        _, ms = run(A, B, mode, out=o, accumulate=acc, split_k=sk, iters=20)
        print("wgrad-shape %s accumulate=%s split%d: %.3f ms  %.0f TF/s" % (
            "pair  " if mode else "single", acc, sk, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
