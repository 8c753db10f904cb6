timeout 600 python -m pytest tests/test_gru_gpu.py -x -q -m gpu > gpurun_out/kt_tests2.txt 2>&1; tail -6 gpurun_out/kt_tests2.txt
for b in 64 32 16 8; do
  TL_B=$b timeout 120 python tools/gru_timeline.py 4 bwd 16 > gpurun_out/tl13_kt_bwd_b$b.txt 2>&1
  TL_B=$b timeout 120 python tools/gru_timeline.py 4 bwd 8 > gpurun_out/tl13_ks_bwd_b$b.txt 2>&1
  echo "B=$b bwd kt: $(grep 'mean step' gpurun_out/tl13_kt_bwd_b$b.txt)  ks: $(grep 'mean step' gpurun_out/tl13_ks_bwd_b$b.txt)"
done
sed -n 1,18p gpurun_out/tl13_kt_bwd_b64.txt
sed -n 1,18p gpurun_out/tl13_kt_bwd_b8.txt
