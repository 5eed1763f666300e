#!/usr/bin/env python
"""Developer probe (GPU box): GCUPS of k_poa_edges on workloads that isolate phases.
  chain : 2 identical reads per edge  -> every row is a fast row, traceback is one diagonal, graph update trivial (fill-bound)
  cfg3  : the bench workload shape (6 noisy reads)
usage: tools/fill_probe.py [n_edges] (HASLR_B200_LIB selects the build)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import haslr_b200  # noqa: E402
import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ctx = haslr_b200.Context(0)
ctx.poa_set_timing(True)
if os.environ.get("HGPU_MAX_WARPS"):
    ctx.poa_configure(0, int(os.environ["HGPU_MAX_WARPS"]))
rng = np.random.default_rng(5)


def tile(bases, seg_off, eso, reps):
    """repeat a small batch `reps` times (generation in numpy is slow)"""
    lens = np.diff(seg_off.astype(np.int64))
    b = np.tile(bases, reps)
    so = np.concatenate(([0], np.cumsum(np.tile(lens, reps)))).astype(np.uint64)
    ne = len(eso) - 1
    per = np.diff(eso.astype(np.int64))
    eo = np.concatenate(([0], np.cumsum(np.tile(per, reps)))).astype(np.uint32)
    return b, so, eo


def run(name, bases, seg_off, eso):
    for _ in range(2):
        ctx.poa_batch(bases, seg_off, eso)
    st = ctx.poa_stats()
    print(f"{name:8s} edges {len(eso)-1:6d} cells {st['cells']/1e9:8.1f} G  kernel {st['ms_dp']:8.1f} ms  {st['cells']/st['ms_dp']/1e6:8.1f} GCUPS", flush=True)


base = 500
# chain: two identical reads
t = synth.ACGT[rng.integers(0, 4, (base, 1500))]
segs = np.repeat(t, 2, axis=0).reshape(-1)
so = (np.arange(2 * base + 1, dtype=np.uint64) * 1500)
eo = (np.arange(base + 1, dtype=np.uint32) * 2)
run("chain", *tile(segs, so, eo, max(1, n // base)))
b, so, eo, _ = synth.poa_batch(11, base, depth=6, length=1500)
run("cfg3", *tile(b, so, eo, max(1, n // base)))
b, so, eo, _ = synth.poa_batch(12, base, depth=2, length=1500)
run("depth2", *tile(b, so, eo, max(1, n // base)))
ctx.close()
