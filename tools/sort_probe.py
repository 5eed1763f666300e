#!/usr/bin/env python
"""Developer probe (GPU box): K2 and K4 on inputs whose per-edge lists are long (a contig end with hundreds to thousands of
supports), the case the rank sorts were quadratic in. Prints the CUDA-event stage times. Build a second library with
-DHGPU_SORT_RANK_MAX=1000000 (tools/build_variant.sh) and set HASLR_B200_LIB to compare against the rank sort."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import haslr_b200  # noqa: E402
import coords_cases  # noqa: E402

ctx = haslr_b200.Context(0)
ctx.set_timing(True)
for max_supp in (60, 700, 3000, 12000):
    c = coords_cases.random_case(11, n_edges=64, max_supp=max_supp, n_reads=2000)
    for _ in range(2):
        ctx.edge_coords(c["edge_rev"], c["supp_off"], c["supp"], c["elems"], c["cl_off"], c["read_len"], c["hits"])
    print(f"K4: 64 edges, up to {max_supp} supports each ({len(c['supp'])} in all): {ctx.stage_stats()['ms_k4']:.3f} ms", flush=True)
rng = np.random.default_rng(5)
for n_contigs, n_reads in ((3000, 50000), (50, 50000), (2, 50000)):
    lens = rng.integers(0, 9, n_reads)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint32)
    tid = rng.integers(0, n_contigs, int(off[-1])).astype(np.uint32)
    rev = rng.integers(0, 2, int(off[-1])).astype(np.uint8)
    for _ in range(2):
        key, soff, supp, keep = ctx.backbone_edges(tid, rev, off, 3)
    n = np.diff(soff)
    print(f"K2: {n_contigs} contigs, {len(supp)} supports over {len(key)} keys (longest list {int(n.max())}): {ctx.stage_stats()['ms_k2']:.3f} ms", flush=True)
