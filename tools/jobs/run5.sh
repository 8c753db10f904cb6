for f in tests/test_conv_gpu.py tests/test_gru_gpu.py tests/test_seq2seq_gpu.py tests/test_models_gpu.py tests/test_zz_configs_gpu.py tests/test_zz_northstar_grads_gpu.py; do
  echo "=== $f"; timeout 600 python -m pytest $f -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pt_$(basename $f .py).txt; tail -4 gpurun_out/pt_$(basename $f .py).txt
done
python tools/gru_timeline.py 4 fwd > gpurun_out/tl5_fwd_ks.txt 2>&1
python tools/gru_timeline.py 4 bwd > gpurun_out/tl5_bwd.txt 2>&1
grep -v "^   P:\|proxy\|epi barrier" gpurun_out/tl5_fwd_ks.txt gpurun_out/tl5_bwd.txt
python tools/debug_step.py > gpurun_out/debug_step.txt 2>&1
tail -34 gpurun_out/debug_step.txt
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','gru_cluster')}, d['e2e']['value'])
    for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step']): print('  %-18s n=%5.1f %.3f ms'%(k,v['launches_per_step'],v['ms_per_step']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
