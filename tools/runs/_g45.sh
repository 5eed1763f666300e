# config 4 scale on one GPU with the staging buffers (2.4 GB of PAF text + 1.9 GB of segments page-locked)
mkdir -p gpurun_out
timeout 1500 python tools/scale_probe.py 14 > gpurun_out/r2_scale14_cfg4.json 2> gpurun_out/r2_scale14_cfg4.err
