# traceback counters + the critical path of lone heavy edges
mkdir -p gpurun_out
L=gpurun_out/r2w_tb.log
: > $L
export HASLR_B200_LIB=build/var/pclk.so
for shape in "592 28 2500" "2368 28 2500" "8 30 3400" "8 6 1500" "148 30 3400"; do
  echo "== $shape" >> $L
  timeout 300 python tools/deep_probe.py $shape 1 2>&1 | grep "phase\|rep 1\|traceback" | tail -3 | cut -c1-400 >> $L
done
