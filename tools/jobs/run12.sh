python - > gpurun_out/kt_tests.txt 2>&1 <<'PY'
import sys, pytest
sys.path.insert(0, ".")
from speech_b200 import _lib
lib = _lib.load()
lib.sb_debug_gru_flags(16)
rc = pytest.main(["tests/test_gru_gpu.py", "-x", "-q", "-m", "gpu"])
print("rc", rc)
PY
tail -8 gpurun_out/kt_tests.txt
for b in 64 32 16 8; do
  TL_B=$b timeout 120 python tools/gru_timeline.py 4 fwd 16 > gpurun_out/tl12_kt_b$b.txt 2>&1
  TL_B=$b timeout 120 python tools/gru_timeline.py 4 fwd 0 > gpurun_out/tl12_ks_b$b.txt 2>&1
  echo "B=$b kt: $(grep 'mean step' gpurun_out/tl12_kt_b$b.txt)  ks: $(grep 'mean step' gpurun_out/tl12_ks_b$b.txt)"
done
sed -n 1,30p gpurun_out/tl12_kt_b64.txt
