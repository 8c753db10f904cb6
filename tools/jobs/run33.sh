( time timeout 900 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/pytest_gpu_all.txt 2>&1; tail -6 gpurun_out/pytest_gpu_all.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
