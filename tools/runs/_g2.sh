mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2b_pytest.log
echo "== A default 592" > gpurun_out/r2b_deep.log
DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2b_deep.log 2>&1
echo "== B team4 x4/SM all edges" >> gpurun_out/r2b_deep.log
HGPU_TEAM=4 HGPU_TEAMS_PER_SM=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2b_deep.log 2>&1
echo "== E default, 2368 edges" >> gpurun_out/r2b_deep.log
timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2b_deep.log 2>&1
(SKIP_REF=1 timeout 600 bash tools/pipeline_cfg2.sh /tmp/cfg2 2>&1 | tail -12) > gpurun_out/r2b_cfg2.log
