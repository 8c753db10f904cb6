python tools/gru_timeline.py 4 fwd > gpurun_out/tl9_fwd_ks.txt 2>&1; grep "mean step\|grid_wait\|barrier seen" gpurun_out/tl9_fwd_ks.txt
python tools/gru_timeline.py 4 bwd > gpurun_out/tl9_bwd.txt 2>&1; grep "mean step" gpurun_out/tl9_bwd.txt
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 300 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'])
print(d.get('dp_check')); print(d.get('secondary'))
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:4]: print('  %-18s n=%5.1f %.3f ms'%(k,v['launches_per_step'],v['ms_per_step']))
PY
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value'], d['ms_per_step'])"
