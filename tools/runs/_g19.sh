# K1 with shared-memory staging: parity tests of the pre-POA stages, then config 2 with the pool's time line
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_k12_gpu.py tests/test_golden_gpu.py tests/test_drop_in_gpu.py tests/test_paf_gpu.py -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r2s_pytest.log
HGPU_VERBOSE=2 PATH_PROBE_STEPS=1 timeout 300 python tools/path_probe.py 2>&1 | grep -v "cleaning" | tail -60 | cut -c1-1200 > gpurun_out/r2s_cfg2_path.log
