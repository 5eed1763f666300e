mkdir -p gpurun_out
echo "== phase clocks, 592 lone warps" > gpurun_out/r2c_phase.log
HASLR_B200_LIB=build/var/pc.so timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2c_phase.log 2>&1
echo "== phase clocks, 2368 lone warps" >> gpurun_out/r2c_phase.log
HASLR_B200_LIB=build/var/pc.so timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2c_phase.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_poa_edges_deep -s 1 -c 1 -o gpurun_out/r2c_deep python tools/deep_probe.py 1184 28 2500 0 > gpurun_out/r2c_ncu.log 2>&1
