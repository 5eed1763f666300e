#!/usr/bin/env python
"""Developer probe (GPU box): deep, long backbone edges as BASELINE config 2 produces them (25x coverage, 2-7 kb gaps),
where a handful of huge score matrices decide the wall time. usage: tools/big_edge_probe.py n_edges depth length [reps]
(HASLR_B200_LIB selects the build, HGPU_VERBOSE=1 prints the pass plan, HGPU_TEAM_MIN_CELLS / HGPU_TEAM steer the team kernel)"""
import os
import sys
import time


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import haslr_b200  # noqa: E402
import synth  # noqa: E402

n, depth, length = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
bases, seg_off, eso, _ = synth.poa_batch(11, n, depth=depth, length=length, length_jitter=0.1)
ctx = haslr_b200.Context(0)
ctx.poa_set_timing(True)
for r in range(reps):
    t0 = time.time()
    cons, off, status = ctx.poa_batch(bases, seg_off, eso)
    dt = time.time() - t0
    st = ctx.poa_stats()
    print(f"edges {n} depth {depth} len {length}: cells {st['cells']/1e9:.1f} G, {st['alignments']} alignments ({st['alignments_i32']} int32), "
          f"kernel {st['ms_dp']:.1f} ms = {st['cells']/st['ms_dp']/1e6:.1f} GCUPS, call {dt*1e3:.0f} ms, status ok {int((status == 0).sum())}/{n}", flush=True)
