mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2p_pytest.log
echo "== cfg3 50k edges: default lib twice" > gpurun_out/r2p_ab.log
bash tools/ab.sh haslr_b200/libhaslr_b200.so haslr_b200/libhaslr_b200.so >> gpurun_out/r2p_ab.log 2>&1
for lib in haslr_b200/libhaslr_b200.so build/var/ctx16.so; do
  echo "== $lib: pool 592 / 2368 / 4736, path" >> gpurun_out/r2p_ab.log
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> gpurun_out/r2p_ab.log
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> gpurun_out/r2p_ab.log
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 4736 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> gpurun_out/r2p_ab.log
  HASLR_B200_LIB=$lib PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0" | tail -1 | cut -c1-220 >> gpurun_out/r2p_ab.log
done
