#!/usr/bin/env python
"""Developer probe (GPU box): the whole path at a multiple of BASELINE config 2 (default x14 = config 4, D. melanogaster scale:
140 Mb genome, 700k reads) through libhaslr_path.so - per-stage wall time and per-kernel CUDA-event time / algorithmic GB/s, i.e.
the K0 / K1 / K2 / K4 roofline entries at a size where those kernels are no longer launch-latency-sized.
usage: tools/scale_probe.py [scale]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import haslr_b200  # noqa: E402
import bench  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 14
bench.CFG2 = dict(genome=10_000_000 * scale, reads=50_000 * scale, read_len=8000, seed=1)
t = time.time()
ctx = haslr_b200.Context(0)
args = argparse.Namespace(steps=1, whole_path_ref=False)
peak, _ = bench.peaks()
out = bench.whole_path_leg(ctx, args, peak)
out["generated_and_run_in_s"] = time.time() - t
print(json.dumps(out, indent=1))
