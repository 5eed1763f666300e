mkdir -p gpurun_out
timeout 1500 python tools/scale_probe.py 14 > gpurun_out/r2m_scale14.json 2> gpurun_out/r2m_scale14.err
tail -3 gpurun_out/r2m_scale14.err
timeout 1200 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -3 gpurun_out/r2m_bench.err
