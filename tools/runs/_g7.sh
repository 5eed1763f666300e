mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2g_pytest.log
echo "== phase clocks, 592 lone warps (HGPU_POOL=0)" > gpurun_out/r2g_phase.log
HGPU_POOL=0 HASLR_B200_LIB=build/var/pc.so timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2g_phase.log 2>&1
echo "== phase clocks, 2368 lone warps (HGPU_POOL=0)" >> gpurun_out/r2g_phase.log
HGPU_POOL=0 HASLR_B200_LIB=build/var/pc.so timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2g_phase.log 2>&1
echo "== pool 592" > gpurun_out/r2g_deep.log
timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2g_deep.log 2>&1
echo "== pool 2368" >> gpurun_out/r2g_deep.log
HGPU_VERBOSE=1 timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2g_deep.log 2>&1
echo "== lone 2368" >> gpurun_out/r2g_deep.log
HGPU_POOL=0 timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2g_deep.log 2>&1
echo "== path (pool)" > gpurun_out/r2g_path.log
HGPU_VERBOSE=1 timeout 300 python tools/path_probe.py 2>&1 | grep -v "wave\|cleaning" | tail -12 >> gpurun_out/r2g_path.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -5 gpurun_out/r2g_bench.err
