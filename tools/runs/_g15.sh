mkdir -p gpurun_out
echo "== A/B shallow kernel: chain-walking sort (default) vs record-based sort; cfg3 50k edges" > gpurun_out/r2o_ab_strec.log
bash tools/ab.sh haslr_b200/libhaslr_b200.so build/var/strec.so haslr_b200/libhaslr_b200.so build/var/strec.so >> gpurun_out/r2o_ab_strec.log 2>&1
echo "== pool 592" > gpurun_out/r2o_deep.log
DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2o_deep.log 2>&1
echo "== pool 2368" >> gpurun_out/r2o_deep.log
timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2o_deep.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2o_pytest.log
