for k in s2s_cell_bwd_kernel s2s_cell_fwd_kernel s2s_attn_fwd_kernel s2s_attn_bwd_b_kernel s2s_attn_bwd_a_kernel; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -s 60 -c 1 -f -o gpurun_out/prof_r02b_$k python tools/s2s_breakdown.py > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
