# pool: first edges dealt by the host (LPT over blocks)
mkdir -p gpurun_out
L=gpurun_out/r2v_ab.log
: > $L
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2v_pytest.log
DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -2 | cut -c1-150 >> $L
timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
HGPU_VERBOSE=2 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|time line\|k_poa_pool:\|first edges\|edge [0-9]*:\|value" | tail -20 | cut -c1-260 >> $L
