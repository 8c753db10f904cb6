"""Regenerate profiles/r01_summary.md from gpurun_out/launches_r01.csv and gpurun_out/prof_r01_*.ncu-rep
(needs `ncu` on PATH for the .ncu-rep import; no GPU required)."""
import collections, csv, io, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []
out.append("# Round %s profile summaries (1x B200, Nsight Compute, `--clock-control none`)\n" % R[1:].lstrip("0"))
out.append("Launch list: `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv "
           "--log-file gpurun_out/launches_%s.csv python bench.py --profile --steps 1 --warmup 1` "
           "(ONE full training step of the north-star config: B=64, T=1000, F=80, conv [[32,5,8,2]]x2, 5x biGRU-1024; "
           "raw list: `profiles/%s_launches.csv`).  Per-launch times under ncu are cold-cache and serialised: compare "
           "SHARES with bench.py's live CUDA-event numbers, not absolutes.\n" % (R, R))
rows = list(csv.reader(open("profiles/%s_launches.csv" % R)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; data = rows[hi + 1:]
kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); mu = hdr.index("Metric Unit")
agg = collections.OrderedDict(); tot = 0.0; mine = 0.0; n_all = 0
for r in data:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    u = r[mu]
    ns = v * 1e3 if u == "us" else (v * 1e6 if u == "ms" else v)
    name = r[kn].split("(")[0][:72]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns; tot += ns; n_all += 1
    if "sb::" in r[kn]:
        mine += ns
out.append("## Launch list of one training step: %d launches, %.2f ms GPU time\n" % (n_all, tot / 1e6))
out.append("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ns / 1e6, 100 * ns / tot))
out.append("\nHand-written kernels (`sb::*`): **%.1f%%** of the GPU time of the step; the rest are torch "
           "elementwise/copy kernels (operand casts, concatenations, zero-fills).  No cuDNN / cuBLAS kernel "
           "appears in the step.\n" % (100 * mine / tot))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed.sum"]
KERNELS = {"r01": ["gru_bwd_ks_kernel", "gru_fwd_kernel", "gemm_bf16_tn_pair_kernel",
                   "gemm_bf16_tn_kernel", "ctc_fwd_bwd_kernel"]}.get(
    R, ["gru_fwd_ks_kernel", "gru_bwd_ks_kernel", "gemm_bf16_tn_pair_kernel", "im2col_kernel",
        "col2im_relu_kernel", "ctc_fwd_bwd_kernel", "sgd_clip_step_kernel", "joint_kernel",
        "joint_reduce_slab_kernel", "rnnt_fwd_bwd_kernel", "rnnt_decode_static_kernel",
        "gru_fwd_kt_kernel", "gru_bwd_kt_kernel", "s2s_cell_fwd_kernel", "s2s_attn_fwd_kernel",
        "s2s_attn_bwd_a_kernel", "s2s_attn_bwd_b_kernel", "s2s_cell_bwd_kernel"])
for k in KERNELS:
    txt = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s_%s.ncu-rep" % (R, k), "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(txt)))
    if len(rr) < 3:
        continue
    h, u, v = rr[0], rr[1], rr[2]
    kname = v[h.index("Kernel Name")] if "Kernel Name" in h else k
    out.append("## `%s` - one launch, `ncu --set full --import-source on -k regex:%s -c 1`\n" % (kname[:90], k))
    if k == "gemm_bf16_tn_kernel" and R == "r01":
        out.append("(capture of the single-CTA 128x256 kernel taken BEFORE the CTA-pair kernel replaced it on "
                   "the large GEMMs; kept for comparison - same shape class, gi of layer 1.)\n")
    if k == "ctc_fwd_bwd_kernel" and R == "r01":
        out.append("(capture from the earlier profiling pass of this round; the kernel has not changed since.)\n")
    out.append("| metric | value |\n|---|---|")
    for w in want:
        if w in h:
            i = h.index(w); out.append("| %s | %s %s |" % (w, v[i], u[i]))
    out.append("")
open("profiles/%s_summary.md" % R, "w").write("\n".join(out))
print("\n".join(out))
