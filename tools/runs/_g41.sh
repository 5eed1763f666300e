# protected edges in the pool's first deal (HGPU_POOL_CHAIN / HGPU_POOL_PENALTY) on config 2 + host-side changes of the whole-path leg
mkdir -p gpurun_out
L=gpurun_out/r2K.log
: > $L
run() { echo "-- $*" >> $L; env "$@" PATH_PROBE_STEPS=2 HGPU_VERBOSE=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|time line\|first edges dealt\|scattered\|gathered\|^{\"value\|^\[poa\]   edge" | cut -c1-330 | tail -14 >> $L; }
run HGPU_POOL_CHAIN=0
run HGPU_POOL_CHAIN=40
run HGPU_POOL_CHAIN=40 HGPU_POOL_PENALTY=1.0
run HGPU_POOL_CHAIN=80 HGPU_POOL_PENALTY=0.7
run HGPU_POOL_CHAIN=0
echo "-- deep probe 592 (all edges alike: the charge must not matter)" >> $L
timeout 200 python tools/deep_probe.py 592 28 2500 2 2>&1 | tail -1 | cut -c1-200 >> $L
echo "-- strong-scaling edge set, chain 0 / 40 / 40 + penalty 1" >> $L
for c in 0 40; do HGPU_POOL_CHAIN=$c timeout 300 python bench.py --steps 1 --warmup 1 --edges 2000 --no-cpu --no-deep --no-whole-path 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['strong_scaling']; print('chain $c', d['ms_per_pass'], d['per_rank_kernel_ms'], d['gcups'])" >> $L 2>&1; done
(timeout 900 python -m pytest tests/test_drop_in_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2) >> $L
