# round-2 evidence run: full GPU tests, the default bench, its launch list under ncu, full captures of the two POA kernels
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2_pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2_clocks.csv &
SMI=$!
timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
kill $SMI
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k[_0-9] -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_launches_bench.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_poa_edges$ -s 3 -c 1 -o gpurun_out/r2_k_poa_edges python bench.py --edges 20000 --steps 1 --warmup 3 --no-cpu --no-deep --no-whole-path --no-strong > gpurun_out/r2_ncu_edges.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_poa_pool -s 1 -c 1 -o gpurun_out/r2_k_poa_pool python tools/deep_probe.py 1184 28 2500 1 > gpurun_out/r2_ncu_pool.log 2>&1
HGPU_VERBOSE=1 timeout 300 python tools/deep_probe.py 592 28 2500 2 > gpurun_out/r2_deep_probe.log 2>&1
timeout 300 python tools/deep_probe.py 2368 28 2500 2 >> gpurun_out/r2_deep_probe.log 2>&1
HGPU_POOL=0 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2_deep_probe.log 2>&1
HGPU_POOL=0 timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2_deep_probe.log 2>&1
HGPU_VERBOSE=2 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep -v "cleaning" | tail -44 | cut -c1-260 > gpurun_out/r2_cfg2_path.log
timeout 1500 python tools/scale_probe.py 14 > gpurun_out/r2_scale14_cfg4.json 2> gpurun_out/r2_scale14_cfg4.err
ls -la gpurun_out | tail -20
