( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/pytest_gpu_all.txt 2>&1; tail -6 gpurun_out/pytest_gpu_all.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 200 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','gru_cluster')}, 'e2e', d['e2e']['value'])
print(d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
print({k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
print(json.dumps(d['secondary']['other_configs']))
PY
