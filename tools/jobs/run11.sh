( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/pytest_gpu_all.txt 2>&1; tail -6 gpurun_out/pytest_gpu_all.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','gru_cluster')}, 'e2e', d['e2e']['value'])
print(d['roofline']['frac'], d['cpu_baseline']['value'])
PY
