python tools/gru_timeline.py 4 fwd > gpurun_out/tl2_fwd4.txt 2>&1
python tools/gru_timeline.py 2 fwd > gpurun_out/tl2_fwd2.txt 2>&1
python tools/gru_timeline.py 1 fwd > gpurun_out/tl2_fwd1.txt 2>&1
python tools/gru_timeline.py 4 fwd 32 > gpurun_out/tl2_fwd_v1.txt 2>&1
python tools/gru_timeline.py 4 bwd > gpurun_out/tl2_bwd.txt 2>&1
cat gpurun_out/tl2_fwd4.txt gpurun_out/tl2_fwd2.txt gpurun_out/tl2_fwd1.txt gpurun_out/tl2_fwd_v1.txt gpurun_out/tl2_bwd.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1
tail -40 gpurun_out/pytest_gpu.txt
python tools/debug_step.py > gpurun_out/debug_step.txt 2>&1
tail -30 gpurun_out/debug_step.txt
