mkdir -p gpurun_out
L=gpurun_out/r2B.log
: > $L
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2B_pytest.log
for v in "" w10; do
  if [ -z "$v" ]; then lib=haslr_b200/libhaslr_b200.so; pl=haslr_b200/libhaslr_path.so; else lib=build/var/$v/libhaslr_b200.so; pl=build/var/$v/libhaslr_path.so; fi
  echo "== $lib" >> $L
  HASLR_B200_LIB=$lib DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -2 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib HASLR_PATH_LIB=$pl HGPU_VERBOSE=2 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|time line\|k_poa_pool:" | tail -3 | cut -c1-260 >> $L
done
