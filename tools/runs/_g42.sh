# protected edges with the tightened threshold + pinned staging of segments / PAF text, on config 2
mkdir -p gpurun_out
L=gpurun_out/r2L.log
: > $L
run() { echo "-- $*" >> $L; env "$@" PATH_PROBE_STEPS=3 HGPU_VERBOSE=2 timeout 300 python tools/path_probe.py > gpurun_out/_pp.txt 2>&1; grep "first edges dealt\|time line" gpurun_out/_pp.txt | tail -2 | cut -c1-250 >> $L; grep "^\[poa\]   edge" gpurun_out/_pp.txt | tail -12 | head -4 | cut -c1-120 >> $L; grep "gpu 0\|gathered\|^{\"value" gpurun_out/_pp.txt | tail -3 | cut -c1-330 >> $L; }
run HGPU_POOL_CHAIN=0
run HGPU_POOL_CHAIN=40
run HGPU_POOL_CHAIN=40 HGPU_POOL_PENALTY=0.8
run HGPU_POOL_CHAIN=40 HGPU_POOL_PROT=0.65
run HGPU_POOL_CHAIN=0
run HGPU_POOL_CHAIN=40
rm -f gpurun_out/_pp.txt
(timeout 900 python -m pytest tests/test_drop_in_gpu.py tests/test_poa_gpu.py -m gpu -x -q 2>&1 | tail -2) >> $L
