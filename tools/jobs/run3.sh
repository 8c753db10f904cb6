for f in tests/test_gru_gpu.py tests/test_gemm_gpu.py tests/test_conv_gpu.py tests/test_joint_gpu.py tests/test_decode_static.py tests/test_rnnt_gpu.py tests/test_models_gpu.py tests/test_zz_configs_gpu.py tests/test_zz_northstar_grads_gpu.py tests/test_ctc_gpu.py tests/test_decode_gpu.py tests/test_specgram_gpu.py tests/test_dp_gpu.py; do
  echo "=== $f"; timeout 600 python -m pytest $f -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pt_$(basename $f .py).txt; tail -4 gpurun_out/pt_$(basename $f .py).txt
done
python tools/gru_timeline.py 4 fwd > gpurun_out/tl3_fwd_ks.txt 2>&1
python tools/gru_timeline.py 4 fwd 32 > gpurun_out/tl3_fwd_v1.txt 2>&1
python tools/gru_timeline.py 4 bwd > gpurun_out/tl3_bwd.txt 2>&1
cat gpurun_out/tl3_fwd_ks.txt gpurun_out/tl3_fwd_v1.txt gpurun_out/tl3_bwd.txt
python tools/debug_step.py > gpurun_out/debug_step.txt 2>&1
tail -12 gpurun_out/debug_step.txt
cat gpurun_out/northstar_grads_bf16.txt 2>/dev/null | head -60
