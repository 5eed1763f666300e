mkdir -p gpurun_out
L=gpurun_out/r2y.log
: > $L
bash tools/ab.sh haslr_b200/libhaslr_b200.so build/var/stbp.so haslr_b200/libhaslr_b200.so build/var/stbp.so >> $L 2>&1
HASLR_B200_LIB=build/var/stbp.so timeout 600 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2 >> $L
for shape in "592 28 2500" "20000 6 1500"; do
  echo "== $shape" >> $L
  HASLR_B200_LIB=build/var/pclk.so timeout 300 python tools/deep_probe.py $shape 1 2>&1 | grep "phase\|rep 1\|traceback\|toposort" | tail -4 | cut -c1-400 >> $L
done
