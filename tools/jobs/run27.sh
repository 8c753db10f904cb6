( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/pytest_gpu_all.txt 2>&1; tail -6 gpurun_out/pytest_gpu_all.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','gru_cluster')}, 'e2e', d['e2e']['value'])
print(d['roofline']['frac'], d['cpu_baseline']['value'])
print(json.dumps(d['secondary']['other_configs'], indent=1))
print({k: d['secondary'][k] for k in d['secondary'] if k.startswith('ctc_loss') or k.startswith('infer')})
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:joint_kernel -s 0 -c 1 -f -o gpurun_out/prof_r02_joint_kernel python tools/profile_other.py > gpurun_out/ncu_joint_kernel.log 2>&1; tail -1 gpurun_out/ncu_joint_kernel.log
