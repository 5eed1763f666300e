# A/B: row bases + plan words of the REL fill in shared memory (build/var/smb.so, -DHGPU_REL_SMEM_BASES=1) vs the batch-register shuffles (default)
mkdir -p gpurun_out
L=gpurun_out/r2J.log
echo "== deep probe 592 / 2368 edges and config 2 K3: default vs smb (x2, interleaved)" > $L
for rep in 1 2; do
for v in default smb; do
  if [ $v = default ]; then export HASLR_B200_LIB=haslr_b200/libhaslr_b200.so HASLR_PATH_LIB=haslr_b200/libhaslr_path.so; else export HASLR_B200_LIB=build/var/smb.so HASLR_PATH_LIB=build/var/smb/libhaslr_path.so; fi
  echo "-- $v" >> $L
  timeout 200 python tools/deep_probe.py 592 28 2500 2 2>&1 | tail -1 | cut -c1-200 >> $L
  timeout 200 python tools/deep_probe.py 2368 28 2500 1 2>&1 | tail -1 | cut -c1-200 >> $L
  PATH_PROBE_STEPS=2 timeout 300 python tools/path_probe.py 2>&1 | grep "gpu 0" | tail -2 | cut -c1-230 >> $L
done
done
HASLR_B200_LIB=build/var/smb.so timeout 600 python -m pytest tests/test_poa_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2 >> $L
