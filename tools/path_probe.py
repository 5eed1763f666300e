#!/usr/bin/env python
"""Developer probe (GPU box): the whole-path leg of bench.py alone (BASELINE config 2 through libhaslr_path.so), optionally after
a cfg3 call in the same context (PATH_PROBE_WARM=n_edges) to see whether earlier calls change the plan. HGPU_VERBOSE=1 prints the plan."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import haslr_b200  # noqa: E402
import bench  # noqa: E402

ctx = haslr_b200.Context(0)
ctx.poa_set_timing(True)
warm = int(os.environ.get("PATH_PROBE_WARM", "0"))
if warm:
    dev = torch.device("cuda", 0)
    d_bases, seg_off, eso = bench.gen_cfg3_torch(warm, 1000, dev)
    d_out = torch.empty(int(seg_off[-1]) // 3 + 4096, dtype=torch.uint8, device=dev)
    for _ in range(2):
        ctx.poa_batch_dev(d_bases.data_ptr(), seg_off, eso, d_out.data_ptr(), d_out.numel(), *bench.SCORES)
    print("warm:", ctx.poa_stats()["ms_dp"], "ms", file=sys.stderr)
args = argparse.Namespace(steps=int(os.environ.get("PATH_PROBE_STEPS", "2")), whole_path_ref=False)
peak, _ = bench.peaks()
out = bench.whole_path_leg(ctx, args, peak)
print(json.dumps({k: out[k] for k in ("value", "s_per_pass", "stage_wall_s")}), file=sys.stderr)
for k in out["kernels"]:
    print(json.dumps({q: k[q] for q in ("kernel", "ms", "launches", "frac") if q in k}), file=sys.stderr)
print(json.dumps(out["kernels"][-1]), file=sys.stderr)
