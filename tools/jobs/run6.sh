for f in tests/test_zz_northstar_grads_gpu.py tests/test_conv_gpu.py; do
  echo "=== $f"; timeout 600 python -m pytest $f -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pt_$(basename $f .py).txt; tail -4 gpurun_out/pt_$(basename $f .py).txt
done
cat gpurun_out/parity_mode.txt
( time python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real
tail -c 600 gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','gru_cluster')}, 'e2e', d['e2e']['value'], d['e2e']['input_pipeline'][:20])
    print('cpu_baseline', d.get('cpu_baseline'))
    sec=d.get('secondary',{})
    for k,v in sec.items():
        if k!='other_configs': print(' ',k,v if not isinstance(v,dict) else json.dumps(v)[:200])
    for k,v in sec.get('other_configs',{}).items(): print('  cfg',k,v)
except Exception as e:
    print('bench parse failed', e)
PY
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
cat gpurun_out/bench_ref.json | cut -c1-900
