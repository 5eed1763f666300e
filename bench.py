#!/usr/bin/env python
"""bench.py — long-read Mbases/s through the batched-POA hot path (BASELINE.json metric) on N B200s.

Workload (config.workload): BASELINE config 3, "batched-POA stress: 200k edges x 6 supporting reads x 1.5 kb gap",
per GPU (weak scaling: every rank owns its own 200k backbone edges, no data-path collective until the final
all-gather of the per-shard consensus). One step = one pass of hgpu_poa_batch over the whole edge batch.

  value      Mbases/s, device-resident inputs (hgpu_poa_batch_dev), CUDA events, max over ranks
  e2e        same metric through the host-buffer C-ABI call hgpu_poa_batch: pinned host bases in, consensus out
  roofline   k_poa_edges: algorithmic bytes (4 B x DP cells + bases in + consensus out) / CUDA-event kernel time
  cpu_baseline  the CPU oracle (restated SPOA 1.1.3, SSE4.1 int16, pthread edge queue) on a bounded sample

`--impl reference` times the reference's CPU path for the same workload: the oracle port of SPOA under the
reference's thread-per-edge queue (real SPOA is an un-vendored dependency of the reference and cannot be built here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "long_read_Mbases_per_s_through_backbone_POA"
UNIT = "Mbases/s"
N_EDGES, DEPTH, GAP_LEN = 200_000, 6, 1500
DEEP_EDGES, DEEP_DEPTH, DEEP_GAP = 592, 28, 2500   # second shape: 4 edges per SM of BASELINE config 2's median edge
ERR = (0.04, 0.03, 0.02)  # ins, del, sub
SCORES = (5, -4, -8)       # Assemble.cpp:8-11


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def gen_cfg3_torch(n_edges, seed, device, chunk=8192, DEPTH=None, GAP_LEN=None):
    """Seeded cfg3 input generated on the GPU: per edge a random 1.5 kb truth and DEPTH noisy copies.
    Returns (bases uint8 cuda tensor, seg_off uint64 numpy, edge_seg_off uint32 numpy)."""
    DEPTH = globals()["DEPTH"] if DEPTH is None else DEPTH
    GAP_LEN = globals()["GAP_LEN"] if GAP_LEN is None else GAP_LEN
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    p_ins, p_del, p_sub = ERR
    parts, lens_all = [], []
    for a in range(0, n_edges, chunk):
        e = min(chunk, n_edges - a)
        truth = torch.randint(0, 4, (e, 1, GAP_LEN), generator=g, device=device, dtype=torch.uint8).expand(e, DEPTH, GAP_LEN)
        truth = truth.reshape(e * DEPTH, GAP_LEN)
        u = torch.rand(truth.shape, generator=g, device=device)
        keep = u >= p_del
        sub = keep & (u < p_del + p_sub)
        shift = torch.randint(1, 4, truth.shape, generator=g, device=device, dtype=torch.uint8)
        code = torch.where(sub, (truth + shift) & 3, truth)
        ins = torch.rand(truth.shape, generator=g, device=device) < p_ins
        ins_code = torch.randint(0, 4, truth.shape, generator=g, device=device, dtype=torch.uint8)
        cnt = keep.to(torch.int64) + ins.to(torch.int64)
        lens = cnt.sum(1)
        row_base = torch.cumsum(lens, 0) - lens
        end = torch.cumsum(cnt, 1) + row_base[:, None]          # exclusive end of each truth position's output
        out = torch.empty(int(lens.sum().item()), dtype=torch.uint8, device=device)
        out[(end - cnt)[keep]] = acgt[code[keep].long()]
        out[(end - 1)[ins]] = acgt[ins_code[ins].long()]
        parts.append(out)
        lens_all.append(lens.cpu().numpy())
        del truth, u, keep, sub, shift, code, ins, ins_code, cnt, end
    bases = torch.cat(parts)
    lens = np.concatenate(lens_all).astype(np.uint64)
    seg_off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
    edge_seg_off = (np.arange(n_edges + 1, dtype=np.uint64) * DEPTH).astype(np.uint32)
    return bases, seg_off, edge_seg_off


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample_edges(cores):
    # ~15 s of oracle work: ~0.5 GCUPS per thread (SSE4.1 int16), ~12.6 Mcells per cfg3 edge
    return int(max(256, min(N_EDGES, 600 * cores)))


def run_oracle_sample(bases_np, seg_off, eso, n, threads):
    import oracle_ffi
    so = seg_off[: int(eso[n]) + 1]
    t = time.perf_counter()
    cons, off, cells, _ = oracle_ffi.poa_batch(bases_np[: int(so[-1])], so, eso[: n + 1], *SCORES, simd=True, threads=threads)
    dt = time.perf_counter() - t
    return cons, off, cells, dt, int(so[-1])


def reference_arm(args):
    """CPU reference arm: the oracle port of the reference's SPOA-per-edge loop on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    cores = os.cpu_count() or 1
    n = cpu_sample_edges(cores) // 4 or 64     # per step; the run does warmup+steps of these
    bases, seg_off, eso, _ = synth.poa_batch(12345, n, depth=DEPTH, length=GAP_LEN, err=ERR)
    for _ in range(args.warmup):
        run_oracle_sample(bases, seg_off, eso, n, cores)
    t = time.perf_counter()
    nb = 0
    for _ in range(args.steps):
        _, _, _, _, b = run_oracle_sample(bases, seg_off, eso, n, cores)
        nb += b
    dt = time.perf_counter() - t
    v = nb / dt / 1e6
    sample = f"{n} of {N_EDGES} cfg3 edges per step ({DEPTH} x {GAP_LEN} bp), restated SPOA 1.1.3 int16 SSE4.1, one edge per queue grab"
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
        "data": "synthetic", "config": {"workload": f"cfg3 batched-POA stress: {N_EDGES} edges x {DEPTH} x {GAP_LEN} bp (bounded sample per step)",
                                        "scores": list(SCORES)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def emit(obj):
    """The one JSON line goes to the real stdout; everything else a library prints (NCCL banners ...) was sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--edges", type=int, default=N_EDGES, help="edges per GPU (default: the cfg3 size; smaller values are for debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-deep", action="store_true", help="skip the deep-edge (config 2 shape) leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import haslr_b200
    from haslr_b200 import sharding

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = haslr_b200.Context(local)
    ctx.poa_set_timing(True)
    if os.environ.get("HGPU_MAX_WARPS"):   # developer knob for occupancy experiments
        ctx.poa_configure(0, int(os.environ["HGPU_MAX_WARPS"]))
    n_edges = args.edges

    # ---- synthetic cfg3 shard of this rank, resident in HBM and mirrored in pinned host memory
    d_bases, seg_off, eso = gen_cfg3_torch(n_edges, 1000 + rank, dev)
    n_bases = int(seg_off[-1])
    h_bases = torch.empty(n_bases, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d_bases)
    out_cap = n_bases // DEPTH * 2 + 4096
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    h_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    def gather_consensus(off):
        """Final exchange of the path: every rank ends up with every shard's consensus (one NCCL all-gather)."""
        if world > 1:
            sharding.all_gather_consensus(dist, d_out, off, dev, to_host=False)

    def step_dev():
        off, status = ctx.poa_batch_dev(d_bases.data_ptr(), seg_off, eso, d_out.data_ptr(), out_cap, *SCORES)
        gather_consensus(off)
        return off, status

    def step_e2e():
        cons, off, status = ctx.poa_batch(h_bases.numpy(), seg_off, eso, *SCORES, out=h_out.numpy())
        return cons, off, status

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also sizes the arena) + correctness spot check against the oracle on rank 0
    for _ in range(max(args.warmup, 3)):
        off, status = step_dev()
    assert (status == 0).all(), f"edge status {np.unique(status)}"
    check = None
    if rank == 0 and not args.no_cpu:
        import oracle_ffi  # checker only
        nchk = min(64, n_edges)
        rc, roff, _, _, _ = run_oracle_sample(h_bases.numpy(), seg_off, eso, nchk, os.cpu_count() or 1)
        got = d_out[: int(off[nchk])].cpu().numpy()
        assert np.array_equal(off[: nchk + 1], roff) and got.tobytes() == rc.tobytes(), "consensus differs from the oracle"
        check = f"first {nchk} edges bit-exact vs oracle"

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    kernel_ms, cells, bases_in, bases_out = 0.0, 0, 0, 0
    barrier()
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        off, status = step_dev()
        st = ctx.poa_stats()
        kernel_ms += st["ms_dp"]; cells += st["cells"]; bases_in += st["bases_in"]; bases_out += st["bases_out"]
        klaunch = st["dp_launches"]
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n_bases * args.steps / (ms_max / 1e3) / 1e6

    # ---- e2e: host buffers through the C ABI, copies inside the timed region
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cons, off, status = step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * n_bases * args.steps / (e2e_ms / 1e3) / 1e6
    h2d = n_bases + seg_off.nbytes + eso.nbytes
    d2h = int(off[-1]) + off.nbytes + status.nbytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peak, peak_src = peaks()
    alg_bytes = 4 * cells + bases_in + bases_out           # 4 B per DP cell + 1 B per base in + consensus out
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "k_poa_edges_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        # ncu cannot replay the 2.3 s / 40 GB launch; DRAM traffic is linear in DP cells, so scale the profiled per-cell figure
        traffic = tj["dram_bytes_per_dp_cell"] * cells / max(1, args.steps * klaunch)
        traffic_src = "profiles/k_poa_edges_traffic.json: %.3f DRAM B per DP cell (ncu, %d-edge launch) x cells of this launch" % (
            tj["dram_bytes_per_dp_cell"], tj["edges_in_profiled_launch"])
    roofline = {"kernel": "k_poa_edges", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "gcups": cells / (kernel_ms / 1e3) / 1e9,
                "kernel_ms_per_launch": kernel_ms / max(1, args.steps * klaunch), "launches_per_step": klaunch,
                "algorithmic_bytes_per_launch": alg_bytes / max(1, args.steps * klaunch)}

    # ---- second shape, reported beside the headline: deep edges as a real 25x dataset produces them (BASELINE config 2:
    #      median 28 supporting reads over a 2.5 kb gap, graphs of ~10^4 nodes, most cells outside the plain int16 range)
    deep = None
    if world == 1 and not args.no_deep:
        dd, dso, deo = gen_cfg3_torch(DEEP_EDGES, 77, dev, chunk=64, DEPTH=DEEP_DEPTH, GAP_LEN=DEEP_GAP)
        dn = int(dso[-1])
        dout = torch.empty(dn // DEEP_DEPTH * 2 + 4096, dtype=torch.uint8, device=dev)
        for _ in range(2):
            doff, dstat = ctx.poa_batch_dev(dd.data_ptr(), dso, deo, dout.data_ptr(), dout.numel(), *SCORES)
        assert (dstat == 0).all(), f"deep edges: status {np.unique(dstat)}"
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(2):
            ctx.poa_batch_dev(dd.data_ptr(), dso, deo, dout.data_ptr(), dout.numel(), *SCORES)
        d1.record(); torch.cuda.synchronize()
        dst = ctx.poa_stats()
        dms = d0.elapsed_time(d1) / 2
        dchk, dcpu = None, None
        if not args.no_cpu:
            rc, roff, ccells, cdt, cnb = run_oracle_sample(dd.cpu().numpy(), dso, deo, 4, os.cpu_count() or 1)
            assert np.array_equal(doff[:5], roff) and dout[: int(doff[4])].cpu().numpy().tobytes() == rc.tobytes(), "deep edges: consensus differs from the oracle"
            dchk = "first 4 edges bit-exact vs oracle"
            # the checker's own speed on those 4 edges (one edge per thread, so 4 host threads busy): context, not a baseline run
            dcpu = {"edges": 4, "threads_busy": 4, "value": cnb / cdt / 1e6, "unit": UNIT, "gcups": ccells / cdt / 1e9}
        deep = {"workload": f"{DEEP_EDGES} edges x {DEEP_DEPTH} supporting reads x {DEEP_GAP} bp gap (BASELINE config 2 edge shape)",
                "value": dn / (dms / 1e3) / 1e6, "unit": UNIT, "ms_per_step": dms, "gcups": dst["cells"] / (dms / 1e3) / 1e9,
                "alignments": dst["alignments"], "alignments_rel16": dst["alignments_rel16"], "alignments_i32": dst["alignments_i32"],
                "kernels": "k_poa_edges_deep (+ k_poa_edges_team for the largest)", "check": dchk, "cpu_oracle_on_check": dcpu}
        del dd, dout

    # ---- CPU baseline: the oracle on a bounded sample of the same edges, all host cores
    cpu = None
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        n = min(cpu_sample_edges(cores), n_edges)
        _, _, ccells, dt, nb = run_oracle_sample(h_bases.numpy(), seg_off, eso, n, cores)
        cpu = {"value": nb / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port", "gcups": ccells / dt / 1e9,
               "sample": f"first {n} of {n_edges} edges of rank 0's shard ({nb / 1e6:.1f} Mbases, {dt:.1f} s), restated SPOA 1.1.3 int16 SSE4.1, {cores} threads"}

    emit(({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": f"cfg3 batched-POA stress: {n_edges} edges x {DEPTH} supporting reads x {GAP_LEN} bp gap per GPU",
                   "edges_per_gpu": n_edges, "segments_per_gpu": int(eso[-1]), "Mbases_per_gpu": n_bases / 1e6, "scores": list(SCORES),
                   "read_error": {"ins": ERR[0], "del": ERR[1], "sub": ERR[2]},
                   "l2": "inputs (1.8 GB) and score matrices (GBs) exceed the 126 MB L2; no flush needed",
                   "sharding": "edges sharded across ranks, NCCL all-gather of consensus at the end" if world > 1 else "single GPU",
                   "check": check},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "deep_edges": deep,
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
