timeout 600 python -m pytest tests/test_seq2seq_gpu.py -x -q -m gpu > gpurun_out/s2s_tests.txt 2>&1; tail -5 gpurun_out/s2s_tests.txt
timeout 600 python -m pytest tests/test_models_gpu.py tests/test_zz_configs_gpu.py -x -q -m gpu > gpurun_out/s2s_tests2.txt 2>&1; tail -5 gpurun_out/s2s_tests2.txt
timeout 300 python tools/s2s_breakdown.py > gpurun_out/s2s_breakdown.txt 2>&1; cat gpurun_out/s2s_breakdown.txt
