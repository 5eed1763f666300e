mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2a_pytest.log
echo "== A default" > gpurun_out/r2a_deep.log
HGPU_VERBOSE=1 timeout 300 python tools/deep_probe.py >> gpurun_out/r2a_deep.log 2>&1
echo "== B team4 x4/SM all edges" >> gpurun_out/r2a_deep.log
HGPU_VERBOSE=1 HGPU_TEAM=4 HGPU_TEAMS_PER_SM=4 timeout 300 python tools/deep_probe.py >> gpurun_out/r2a_deep.log 2>&1
echo "== C team8 x2/SM all edges" >> gpurun_out/r2a_deep.log
HGPU_VERBOSE=1 HGPU_TEAM=8 HGPU_TEAM_ALPHA=0.01 HGPU_TEAM_MIN_CELLS=1 timeout 300 python tools/deep_probe.py >> gpurun_out/r2a_deep.log 2>&1
echo "== D team4 x4/SM, 2368 edges" >> gpurun_out/r2a_deep.log
HGPU_TEAM=4 HGPU_TEAMS_PER_SM=4 HGPU_TEAM_ALPHA=0.01 HGPU_TEAM_MIN_CELLS=1 timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2a_deep.log 2>&1
echo "== E default, 2368 edges" >> gpurun_out/r2a_deep.log
timeout 300 python tools/deep_probe.py 2368 28 2500 1 >> gpurun_out/r2a_deep.log 2>&1
(SKIP_REF= timeout 600 bash tools/pipeline_cfg2.sh /tmp/cfg2 2>&1 | tail -40) > gpurun_out/r2a_cfg2.log
