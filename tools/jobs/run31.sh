NCU="ncu --clock-control none"
$NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r02.csv python bench.py --profile --steps 1 --warmup 1 > gpurun_out/prof_launch.log 2>&1
tail -2 gpurun_out/prof_launch.log; wc -l gpurun_out/launches_r02.csv
$NCU --set full --import-source on --profile-from-start off -k regex:col2im_relu -c 1 -f -o gpurun_out/prof_r02_col2im_relu_kernel python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_col2im.log 2>&1; tail -1 gpurun_out/ncu_col2im.log
