mkdir -p gpurun_out
L=gpurun_out/r2D_sort.log
: > $L
(timeout 900 python -m pytest tests/test_k12_gpu.py tests/test_coords_gpu.py tests/test_drop_in_gpu.py -m gpu -x -q 2>&1 | tail -4) >> $L
echo "== bitonic above 64 (default build)" >> $L
timeout 600 python tools/sort_probe.py >> $L 2>&1
echo "== rank sort only (-DHGPU_SORT_RANK_MAX=100000000)" >> $L
HASLR_B200_LIB=build/var/ranksort.so timeout 900 python tools/sort_probe.py >> $L 2>&1
