mkdir -p gpurun_out
L=gpurun_out/r2E_w16.log
: > $L
for v in w16; do
  lib=build/var/$v/libhaslr_b200.so; pl=build/var/$v/libhaslr_path.so
  echo "== $lib" >> $L
  (HASLR_B200_LIB=$lib timeout 600 python -m pytest tests/test_poa_gpu.py -m gpu -x -q 2>&1 | tail -2) >> $L
  HASLR_B200_LIB=$lib DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -2 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 2368 28 2500 2 2>&1 | tail -2 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib HASLR_PATH_LIB=$pl HGPU_VERBOSE=2 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|time line\|k_poa_pool:\|edge [0-9]*:" | tail -8 | cut -c1-260 >> $L
done
