timeout 600 python -m pytest tests/test_joint_gpu.py tests/test_rnnt_gpu.py -x -q -m gpu > gpurun_out/joint_tests.txt 2>&1; tail -5 gpurun_out/joint_tests.txt
timeout 300 python tools/step_breakdown.py rnnt > gpurun_out/rnnt_breakdown.txt 2>&1; cat gpurun_out/rnnt_breakdown.txt
