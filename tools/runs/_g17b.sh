mkdir -p gpurun_out
HGPU_VERBOSE=1 timeout 600 python bench.py --edges 50000 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
grep -v "^\[poa\]" gpurun_out/r2q_bench.err | tail -30
