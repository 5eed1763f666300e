mkdir -p gpurun_out
HGPU_POOL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_poa_edges_deep -s 1 -c 1 -o gpurun_out/r2j_deep python tools/deep_probe.py 1184 28 2500 1 > gpurun_out/r2j_ncu.log 2>&1
echo "== path (pool)" > gpurun_out/r2j_path.log
HGPU_VERBOSE=1 PATH_PROBE_STEPS=1 timeout 300 python tools/path_probe.py 2>&1 | grep -v "cleaning" | tail -30 | cut -c1-250 >> gpurun_out/r2j_path.log
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2j_pytest.log
