timeout 600 python -m pytest tests/test_joint_gpu.py tests/test_rnnt_gpu.py tests/test_decode_static.py -x -q -m gpu > gpurun_out/joint_tests.txt 2>&1; tail -3 gpurun_out/joint_tests.txt
timeout 300 python tools/step_breakdown.py rnnt > gpurun_out/rnnt_breakdown4.txt 2>&1; head -9 gpurun_out/rnnt_breakdown4.txt
