for fl in 0 64 128; do
  python tools/gru_timeline.py 4 fwd $fl > gpurun_out/tl10_fwd_$fl.txt 2>&1; echo "fwd flags $fl: $(grep 'mean step' gpurun_out/tl10_fwd_$fl.txt) $(grep -m1 'grid_wait' gpurun_out/tl10_fwd_$fl.txt)"
  python tools/gru_timeline.py 4 bwd $fl > gpurun_out/tl10_bwd_$fl.txt 2>&1; echo "bwd flags $fl: $(grep 'mean step' gpurun_out/tl10_bwd_$fl.txt)"
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for f in tests/test_*gpu*.py tests/test_decode_static.py; do
  timeout 900 python -m pytest $f -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pt_$(basename $f .py).txt; echo "$f: $(tail -1 gpurun_out/pt_$(basename $f .py).txt)"
done
