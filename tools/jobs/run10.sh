for b in 8 16 32; do
  for fl in 0 32; do
    TL_B=$b python tools/gru_timeline.py 4 fwd $fl > gpurun_out/tl11_fwd_b${b}_$fl.txt 2>&1; echo "B=$b fwd flags $fl: $(grep 'mean step' gpurun_out/tl11_fwd_b${b}_$fl.txt) cluster $(grep 'cluster size used' gpurun_out/tl11_fwd_b${b}_$fl.txt)"
  done
  TL_B=$b python tools/gru_timeline.py 4 bwd 0 > gpurun_out/tl11_bwd_b${b}.txt 2>&1; echo "B=$b bwd ks: $(grep 'mean step' gpurun_out/tl11_bwd_b${b}.txt)"
  TL_B=$b SB_GRU_KSPLIT=0 python tools/gru_timeline.py 4 bwd 0 > gpurun_out/tl11_bwd_b${b}_plain.txt 2>&1; echo "B=$b bwd plain: $(grep 'mean step' gpurun_out/tl11_bwd_b${b}_plain.txt)"
done
