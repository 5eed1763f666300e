# the one shot left in the round (2.9 GPU-minutes): w_toposort_claims. In-tree build = claims in the deep kernels; build/var/tcs = in the shallow kernel too.
mkdir -p gpurun_out
L=gpurun_out/r2N.log
echo "== parity, in-tree (claims in k_poa_pool / deep / team)" > $L
(timeout 75 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2) >> $L
echo "== deep probe 592 x 28 x 2500, in-tree (before: 200.7-206 ms)" >> $L
timeout 40 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -1 | cut -c1-200 >> $L
echo "== cfg3 50k edges: tcs (claims in the shallow kernel too, with the oracle spot check) vs in-tree" >> $L
HASLR_B200_LIB=build/var/tcs.so timeout 60 python bench.py --edges 50000 --steps 2 --warmup 3 --no-deep --no-whole-path --no-strong 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tcs gcups', round(d['roofline']['gcups'],1), 'value', round(d['value'],1), 'check', d['config'].get('check'))" >> $L 2>&1
timeout 40 python bench.py --edges 50000 --steps 2 --warmup 3 --no-cpu --no-deep --no-whole-path --no-strong 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('in-tree gcups', round(d['roofline']['gcups'],1), 'value', round(d['value'],1))" >> $L 2>&1
echo "== config 2 path, in-tree (before: K3 333-338 ms)" >> $L
PATH_PROBE_STEPS=2 timeout 40 python tools/path_probe.py 2>&1 | grep "gpu 0\|^{\"value" | tail -3 | cut -c1-250 >> $L
