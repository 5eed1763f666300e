# pool claim policies: 0 longest remaining first, 1 graph tasks first then longest remaining, 2 coarse priority, 3 fixed tie-break
mkdir -p gpurun_out
L=gpurun_out/r2u_ab.log
: > $L
for v in "" pol1 pol2 pol3; do
  if [ -z "$v" ]; then lib=haslr_b200/libhaslr_b200.so; pl=haslr_b200/libhaslr_path.so; else lib=build/var/$v/libhaslr_b200.so; pl=build/var/$v/libhaslr_path.so; fi
  echo "== $lib" >> $L
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib HASLR_PATH_LIB=$pl HGPU_VERBOSE=2 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|time line\|k_poa_pool:" | tail -3 | cut -c1-260 >> $L
done
echo "== phase clocks of the pool, 592 and 2368 edges" >> $L
HASLR_B200_LIB=build/var/pclk.so timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | grep "phase\|rep" | cut -c1-400 >> $L
HASLR_B200_LIB=build/var/pclk.so timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | grep "phase\|rep" | cut -c1-400 >> $L
