#!/usr/bin/env python
"""Developer probe (GPU box): bench.py's deep-edge leg alone (edges x reads x gap of BASELINE config 2's median edge),
device-resident, so scheduling variants can be compared through the HGPU_* knobs.
usage: tools/deep_probe.py [n_edges depth gap reps]   (HGPU_VERBOSE=1 prints the pass plan)"""
import os
import sys

import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import haslr_b200  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.DEEP_EDGES
depth = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEEP_DEPTH
gap = int(sys.argv[3]) if len(sys.argv) > 3 else bench.DEEP_GAP
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda", 0)
ctx = haslr_b200.Context(0)
ctx.poa_set_timing(True)
dd, dso, deo = bench.gen_cfg3_torch(n, 77, dev, chunk=64, DEPTH=depth, GAP_LEN=gap)
dout = torch.empty(int(dso[-1]) // depth * 2 + 4096, dtype=torch.uint8, device=dev)
ref = None
for r in range(reps + 1):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    off, status = ctx.poa_batch_dev(dd.data_ptr(), dso, deo, dout.data_ptr(), dout.numel(), *bench.SCORES)
    ev1.record(); torch.cuda.synchronize()
    st = ctx.poa_stats()
    ms = ev0.elapsed_time(ev1)
    h = zlib.crc32(dout[: int(off[-1])].cpu().numpy().tobytes())
    print(f"rep {r}: {n} x {depth} x {gap}: call {ms:.1f} ms, kernels {st['ms_dp']:.1f} ms, {st['cells'] / ms / 1e6:.1f} GCUPS (call), "
          f"{int(dso[-1]) / ms / 1e3:.1f} Mbases/s, rel16 {st['alignments_rel16']}/{st['alignments']}, launches {st['dp_launches']}, "
          f"ok {int((status == 0).sum())}/{n}, cons {int(off[-1])} bytes hash {h & 0xFFFFFFFF:08x}", flush=True)
if os.environ.get("DEEP_PROBE_CHECK"):
    import oracle_ffi
    k = int(os.environ["DEEP_PROBE_CHECK"])
    rc, roff, _, _ = oracle_ffi.poa_batch(dd.cpu().numpy()[: int(dso[int(deo[k])])], dso[: int(deo[k]) + 1], deo[: k + 1], *bench.SCORES, simd=True, threads=os.cpu_count())
    ok = np.array_equal(off[: k + 1], roff) and dout[: int(off[k])].cpu().numpy().tobytes() == rc.tobytes()
    print(f"oracle check on the first {k} edges: {'bit-exact' if ok else 'MISMATCH'}")
