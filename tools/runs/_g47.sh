# final commit: all GPU tests + smoke, then the protected-edge constants once more on config 2
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/r2_pytest_gpu.log
(python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) >> gpurun_out/r2_pytest_gpu.log
L=gpurun_out/r2M.log
: > $L
run() { echo "-- $*" >> $L; env "$@" PATH_PROBE_STEPS=2 HGPU_VERBOSE=2 timeout 200 python tools/path_probe.py > gpurun_out/_pp.txt 2>&1; grep "first edges dealt\|time line" gpurun_out/_pp.txt | tail -2 | cut -c1-250 >> $L; grep "^\[poa\]   edge" gpurun_out/_pp.txt | tail -12 | head -3 | cut -c1-120 >> $L; grep "gpu 0" gpurun_out/_pp.txt | tail -2 | cut -c1-200 >> $L; }
run HGPU_POOL_PENALTY=0.8
run HGPU_POOL_PENALTY=1.3
run HGPU_POOL_PENALTY=1.3 HGPU_POOL_PROT=0.55
run HGPU_POOL_PENALTY=2.0 HGPU_POOL_PROT=0.72
rm -f gpurun_out/_pp.txt
