mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2e_pytest.log
echo "== A default 592" > gpurun_out/r2e_deep.log
DEEP_PROBE_CHECK=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2e_deep.log 2>&1
echo "== B team4 x4/SM all edges" >> gpurun_out/r2e_deep.log
HGPU_TEAM=4 HGPU_TEAMS_PER_SM=4 timeout 300 python tools/deep_probe.py 592 28 2500 1 >> gpurun_out/r2e_deep.log 2>&1
echo "== E default, 2368 edges" >> gpurun_out/r2e_deep.log
timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2e_deep.log 2>&1
(timeout 600 bash tools/pipeline_cfg2.sh /tmp/cfg2 2>&1 | tail -32) > gpurun_out/r2e_cfg2.log
(HGPU_TEAM=4 HGPU_TEAMS_PER_SM=4 HGPU_TEAM_ALPHA=0.3 SKIP_REF=1 timeout 600 bash tools/pipeline_cfg2.sh /tmp/cfg2 2>&1 | tail -14) > gpurun_out/r2e_cfg2_team4.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -5 gpurun_out/r2e_bench.err
