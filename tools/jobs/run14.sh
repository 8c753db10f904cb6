timeout 600 python -m pytest tests/test_seq2seq_gpu.py -x -q -m gpu > gpurun_out/s2s_tests.txt 2>&1; tail -15 gpurun_out/s2s_tests.txt
timeout 600 python -m pytest tests/test_models_gpu.py -x -q -m gpu > gpurun_out/s2s_tests2.txt 2>&1; tail -8 gpurun_out/s2s_tests2.txt
timeout 600 python -m pytest tests/test_zz_configs_gpu.py -x -q -m gpu > gpurun_out/s2s_tests3.txt 2>&1; tail -8 gpurun_out/s2s_tests3.txt
timeout 600 python - > gpurun_out/s2s_other.txt 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
import bench
from speech_b200 import ops
r = bench.other_config_measurements(torch.device("cuda"))
print(json.dumps(r, indent=1))
PY
tail -30 gpurun_out/s2s_other.txt
