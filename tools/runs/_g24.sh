# traceback with predecessor words + first-predecessor chains
mkdir -p gpurun_out
L=gpurun_out/r2x_tb.log
: > $L
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2x_pytest.log
DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -2 | cut -c1-150 >> $L
timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
HGPU_VERBOSE=2 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|time line\|value" | tail -4 | cut -c1-260 >> $L
for shape in "592 28 2500" "8 30 3400"; do
  echo "== $shape" >> $L
  HASLR_B200_LIB=build/var/pclk.so timeout 300 python tools/deep_probe.py $shape 1 2>&1 | grep "phase\|rep 1\|traceback" | tail -3 | cut -c1-400 >> $L
done
