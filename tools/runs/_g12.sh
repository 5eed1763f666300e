mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2l_pytest.log
echo "== pool 592" > gpurun_out/r2l_deep.log
HGPU_VERBOSE=1 DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | grep -v "wave [0-9]*:" >> gpurun_out/r2l_deep.log
echo "== pool 2368" >> gpurun_out/r2l_deep.log
timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2l_deep.log 2>&1
echo "== path (pool, one kernel)" > gpurun_out/r2l_path.log
HGPU_VERBOSE=1 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep -v "cleaning" | tail -22 | cut -c1-250 >> gpurun_out/r2l_path.log
HGPU_VERBOSE=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-deep --no-whole-path > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
grep "pool\|host" gpurun_out/r2l_bench.err | tail -12 | cut -c1-220 > gpurun_out/r2l_bench_plan.log
