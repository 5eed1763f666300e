# headline refresh after the shallow fill's direct boundary store: bench, reference arm, launch list, ncu capture of k_poa_edges
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2_clocks.csv &
SMI=$!
timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
kill $SMI
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k[_0-9] -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_launches_bench.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_poa_edges$ -s 3 -c 1 -f -o gpurun_out/r2_k_poa_edges python bench.py --edges 20000 --steps 1 --warmup 3 --no-cpu --no-deep --no-whole-path --no-strong > gpurun_out/r2_ncu_edges.log 2>&1
(timeout 900 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py tests/test_drop_in_gpu.py -m gpu -q 2>&1 | tail -3) > gpurun_out/r2_pytest_poa_after_bco.log
