python tools/gru_timeline.py 4 fwd > gpurun_out/tl7_fwd_ks.txt 2>&1; tail -8 gpurun_out/tl7_fwd_ks.txt
NCU="ncu --clock-control none"
$NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r02.csv python bench.py --profile --steps 1 --warmup 1 > gpurun_out/prof_launch.log 2>&1
for k in gru_fwd_ks_kernel gru_bwd_ks_kernel gemm_bf16_tn_pair_kernel im2col_kernel col2im_relu_kernel ctc_fwd_bwd_kernel sgd_clip_step_kernel; do
  $NCU --set full --import-source on --profile-from-start off -k regex:$k -c 1 -f -o gpurun_out/prof_r02_$k python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_$k.log 2>&1
  echo "$k: $(tail -1 gpurun_out/ncu_$k.log | cut -c1-120)"
done
for k in joint_kernel rnnt_fwd_bwd_kernel rnnt_decode_static_kernel s2s_attn_fwd_kernel s2s_attn_bwd_kernel s2s_cell_fwd_kernel s2s_cell_bwd_kernel; do
  $NCU --set full --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r02_$k python tools/profile_other.py > gpurun_out/ncu_$k.log 2>&1
  echo "$k: $(tail -1 gpurun_out/ncu_$k.log | cut -c1-120)"
done
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
