mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2f_pytest.log
echo "== pool 592" > gpurun_out/r2f_deep.log
HGPU_VERBOSE=1 DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 2 >> gpurun_out/r2f_deep.log 2>&1
echo "== pool 592 ctx 3" >> gpurun_out/r2f_deep.log
HGPU_POOL_CTX=3 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2f_deep.log 2>&1
echo "== pool 592 ctx 4" >> gpurun_out/r2f_deep.log
HGPU_POOL_CTX=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2f_deep.log 2>&1
echo "== pool 2368" >> gpurun_out/r2f_deep.log
HGPU_VERBOSE=1 timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2f_deep.log 2>&1
echo "== pool 2368 ctx 8" >> gpurun_out/r2f_deep.log
HGPU_POOL_CTX=8 timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2f_deep.log 2>&1
echo "== path fresh (pool)" > gpurun_out/r2f_path.log
HGPU_VERBOSE=1 timeout 300 python tools/path_probe.py >> gpurun_out/r2f_path.log 2>&1
echo "== path fresh (no pool)" >> gpurun_out/r2f_path.log
HGPU_POOL=0 HGPU_VERBOSE=1 timeout 300 python tools/path_probe.py >> gpurun_out/r2f_path.log 2>&1
echo "== path after cfg3 warm (no pool)" >> gpurun_out/r2f_path.log
HGPU_POOL=0 PATH_PROBE_WARM=100000 HGPU_VERBOSE=1 timeout 300 python tools/path_probe.py >> gpurun_out/r2f_path.log 2>&1
