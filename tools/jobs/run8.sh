timeout 600 python -m pytest tests/test_gru_gpu.py -m gpu -q -x 2>&1 | tail -3
python tools/gru_timeline.py 4 fwd > gpurun_out/tl8_fwd_ks.txt 2>&1; grep -v "proxy\|epi barrier" gpurun_out/tl8_fwd_ks.txt | tail -22
python tools/gru_timeline.py 4 bwd > gpurun_out/tl8_bwd.txt 2>&1; tail -3 gpurun_out/tl8_bwd.txt
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','gru_cluster')}, d['e2e']['value'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:4]: print('  %-18s n=%5.1f %.3f ms'%(k,v['launches_per_step'],v['ms_per_step']))
PY
for k in rnnt_fwd_bwd_kernel rnnt_decode_static_kernel; do
  ncu --clock-control none --set full --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_r02_$k python tools/profile_other.py > gpurun_out/ncu_$k.log 2>&1
  echo "$k: $(tail -1 gpurun_out/ncu_$k.log | cut -c1-120)"
done
