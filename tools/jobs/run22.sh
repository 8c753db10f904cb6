python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 300 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gru_cluster')})
print('dp_check', d.get('dp_check'))
print('weak', d.get('secondary',{}).get('weak_scaling'))
print('rnnt', d.get('secondary',{}).get('rnnt_dp'))
PY
timeout 600 python -m pytest tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -3
