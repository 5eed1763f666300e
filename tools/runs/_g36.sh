mkdir -p gpurun_out
L=gpurun_out/r2H.log
: > $L
(timeout 900 python -m pytest tests/test_k12_gpu.py tests/test_golden_gpu.py tests/test_drop_in_gpu.py tests/test_paf_gpu.py -m gpu -x -q 2>&1 | tail -4) >> $L
PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|kernel\|value" | tail -8 | cut -c1-230 >> $L
timeout 1500 python tools/scale_probe.py 14 2>&1 | grep -A3 '"kernel"' | grep "kernel\|\"ms\"" | paste - - | cut -c1-160 >> $L
