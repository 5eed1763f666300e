mkdir -p gpurun_out
echo "== cfg3 50k edges: default, direct boundary store in the shallow fill, x2" > gpurun_out/r2I.log
bash tools/ab.sh haslr_b200/libhaslr_b200.so build/var/f16bco.so haslr_b200/libhaslr_b200.so build/var/f16bco.so >> gpurun_out/r2I.log 2>&1
HASLR_B200_LIB=build/var/f16bco.so HGPU_POOL=0 timeout 600 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/r2I.log
