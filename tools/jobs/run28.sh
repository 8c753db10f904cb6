timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_zz_northstar_grads_gpu.py -x -q -m gpu > gpurun_out/t28.txt 2>&1; tail -3 gpurun_out/t28.txt
python bench.py --gpus 1 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench28.json 2> gpurun_out/bench28.err; tail -c 300 gpurun_out/bench28.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench28.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
for k,v in d['kernels'].items(): print(k, round(v['ms_per_step'],3))
PY
