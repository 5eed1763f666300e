mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py tests/test_drop_in_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2n_pytest.log
echo "== pool 592" > gpurun_out/r2n_deep.log
DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 >> gpurun_out/r2n_deep.log
echo "== pool 2368" >> gpurun_out/r2n_deep.log
timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2n_deep.log 2>&1
echo "== phase clocks, 592 lone warps (HGPU_POOL=0)" >> gpurun_out/r2n_deep.log
HGPU_POOL=0 HASLR_B200_LIB=build/var/pc.so timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2n_deep.log 2>&1
echo "== A/B traceback tile copies: LDG+STS (default) vs cp.async (LDGSTS); cfg3 50k edges, then deep 2368" > gpurun_out/r2n_ab_cpasync.log
bash tools/ab.sh haslr_b200/libhaslr_b200.so build/var/cpasync.so haslr_b200/libhaslr_b200.so build/var/cpasync.so >> gpurun_out/r2n_ab_cpasync.log 2>&1
for lib in haslr_b200/libhaslr_b200.so build/var/cpasync.so; do echo $lib >> gpurun_out/r2n_ab_cpasync.log; HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 >> gpurun_out/r2n_ab_cpasync.log; done
echo "== path" > gpurun_out/r2n_path.log
PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep -v "cleaning" | tail -6 | cut -c1-300 >> gpurun_out/r2n_path.log
