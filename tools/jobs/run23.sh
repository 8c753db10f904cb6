# refresh of the per-kernel captures for the kernels rewritten late in round 2
for k in joint_reduce_slab_kernel gru_fwd_kt_kernel gru_bwd_kt_kernel s2s_cell_fwd_kernel s2s_attn_fwd_kernel s2s_attn_bwd_a_kernel s2s_attn_bwd_b_kernel s2s_cell_bwd_kernel; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_r02_$k python tools/profile_other.py > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
rm -f gpurun_out/prof_r02_s2s_attn_bwd_kernel.ncu-rep
