mkdir -p gpurun_out
L=gpurun_out/r2F.log
: > $L
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) >> $L
PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|kernel\|value" | tail -9 | cut -c1-260 >> $L
