mkdir -p gpurun_out
echo "== cfg3 50k edges: default lib twice" > gpurun_out/r2q_ab.log
bash tools/ab.sh haslr_b200/libhaslr_b200.so haslr_b200/libhaslr_b200.so >> gpurun_out/r2q_ab.log 2>&1
for lib in haslr_b200/libhaslr_b200.so; do
  echo "== $lib: pool 592 / 2368, path" >> gpurun_out/r2q_ab.log
  HASLR_B200_LIB=$lib DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -2 | cut -c1-150 >> gpurun_out/r2q_ab.log
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> gpurun_out/r2q_ab.log
  HASLR_B200_LIB=$lib PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|value" | tail -2 | cut -c1-220 >> gpurun_out/r2q_ab.log
done
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2q_pytest.log
