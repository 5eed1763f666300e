mkdir -p gpurun_out
echo "== path (pool), after the hardware-queue fix" > gpurun_out/r2k_path.log
HGPU_VERBOSE=1 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep -v "cleaning\|^\[poa\] attempt" | tail -16 | cut -c1-250 >> gpurun_out/r2k_path.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
tail -5 gpurun_out/r2k_bench_n2.err
