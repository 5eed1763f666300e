mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_poa_edges_deep -s 1 -c 1 -o gpurun_out/r2d_deep python tools/deep_probe.py 1184 28 2500 1 > gpurun_out/r2d_ncu.log 2>&1
ls -la gpurun_out/ >> gpurun_out/r2d_ncu.log
