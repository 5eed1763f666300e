mkdir -p gpurun_out
L=gpurun_out/r2C.log
: > $L
for v in "" fscan bco bcofs uni; do
  if [ -z "$v" ]; then lib=haslr_b200/libhaslr_b200.so; else lib=build/var/$v/libhaslr_b200.so; fi
  echo "== $lib" >> $L
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
  HASLR_B200_LIB=$lib timeout 300 python tools/deep_probe.py 2368 28 2500 2 2>&1 | tail -2 | cut -c1-150 >> $L
done
