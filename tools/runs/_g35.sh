mkdir -p gpurun_out
L=gpurun_out/r2G.log
: > $L
run() {   # label, lib dir or "", env
  echo "== $1" >> $L
  if [ -z "$2" ]; then lib=haslr_b200/libhaslr_b200.so; pl=haslr_b200/libhaslr_path.so; else lib=build/var/$2/libhaslr_b200.so; pl=build/var/$2/libhaslr_path.so; fi
  env $3 HASLR_B200_LIB=$lib HASLR_PATH_LIB=$pl HGPU_VERBOSE=2 PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0\|time line\|dedicated\|k1_compact" | tail -5 | cut -c1-230 >> $L
}
run "main: 1 x 16 warps, dedicated warps alpha 0.6, K1 scratch in smem" "" "A=1"
run "main, no dedicated warps" "" "HGPU_CRIT_ALPHA=1000"
run "main, alpha 0.45" "" "HGPU_CRIT_ALPHA=0.45"
run "main, alpha 0.8" "" "HGPU_CRIT_ALPHA=0.8"
run "2 x 8 warps, 16 contexts, K1 scratch global" "w8" "HGPU_CRIT_ALPHA=1000"
run "main again" "" "A=1"
for a in 0.6 1000; do
  echo "== strong-leg-like: deep probe 2368 and 592, alpha $a" >> $L
  HGPU_CRIT_ALPHA=$a timeout 300 python tools/deep_probe.py 592 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
  HGPU_CRIT_ALPHA=$a timeout 300 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-150 >> $L
done
