python tools/mn_probe.py > gpurun_out/mn_probe.txt 2>&1
python tools/gru_timeline.py 8 fwd > gpurun_out/tl_fwd8.txt 2>&1
python tools/gru_timeline.py 4 fwd > gpurun_out/tl_fwd4.txt 2>&1
python tools/gru_timeline.py 1 fwd > gpurun_out/tl_fwd1.txt 2>&1
python tools/gru_timeline.py 8 bwd > gpurun_out/tl_bwd.txt 2>&1
python tools/gru_ablate.py 0 1 2 3 > gpurun_out/ablate.txt 2>&1
cat gpurun_out/mn_probe.txt gpurun_out/tl_fwd8.txt gpurun_out/tl_fwd4.txt gpurun_out/tl_fwd1.txt gpurun_out/tl_bwd.txt gpurun_out/ablate.txt
