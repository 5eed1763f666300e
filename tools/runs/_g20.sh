# pool claims by longest remaining critical path; 3 blocks per SM with a ring of 4 parked rows
mkdir -p gpurun_out
L=gpurun_out/r2t_ab.log
: > $L
for lib in haslr_b200/libhaslr_b200.so build/var/b2r4.so build/var/b3r4.so; do
  echo "== $lib: pool 592 / 2368, path" >> $L
  HASLR_B200_LIB=$lib DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -2 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib HGPU_VERBOSE=2 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|value\|time line\|k_poa_pool:" | tail -5 | cut -c1-260 >> $L
done
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2t_pytest.log
