mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2h_pytest.log
echo "== pool 592" > gpurun_out/r2h_deep.log
DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2h_deep.log 2>&1
echo "== pool 2368" >> gpurun_out/r2h_deep.log
timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2h_deep.log 2>&1
echo "== path (pool)" > gpurun_out/r2h_path.log
HGPU_VERBOSE=1 timeout 300 python tools/path_probe.py 2>&1 | grep -v "wave [0-9]*:\|cleaning" | tail -24 | cut -c1-250 >> gpurun_out/r2h_path.log
HGPU_VERBOSE=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-deep --no-whole-path > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
grep -v "wave [0-9]*:" gpurun_out/r2h_bench.err | tail -60 | cut -c1-220 > gpurun_out/r2h_bench_plan.log
