python tools/mn_probe.py > gpurun_out/mn_probe.txt 2>&1
tail -8 gpurun_out/mn_probe.txt
python tools/gru_timeline.py 8 fwd > gpurun_out/tl_fwd8.txt 2>&1
python tools/gru_timeline.py 4 fwd > gpurun_out/tl_fwd4.txt 2>&1
python tools/gru_timeline.py 1 fwd > gpurun_out/tl_fwd1.txt 2>&1
python tools/gru_timeline.py 8 bwd > gpurun_out/tl_bwd.txt 2>&1
python tools/gru_ablate.py 0 1 2 3 > gpurun_out/ablate.txt 2>&1
cat gpurun_out/tl_fwd8.txt gpurun_out/tl_fwd4.txt gpurun_out/tl_fwd1.txt gpurun_out/tl_bwd.txt gpurun_out/ablate.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
tail -15 gpurun_out/pytest_gpu.txt
python tools/debug_step.py > gpurun_out/debug_step.txt 2>&1
tail -40 gpurun_out/debug_step.txt
