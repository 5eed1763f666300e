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


def gen_cfg3_torch(n_edges, seed, device, chunk=4096, DEPTH=None, GAP_LEN=None):
    """Seeded cfg3 input generated on the GPU by tests/synth.hashed_batch (counter-based: the numpy backend of the same function
    gives the CPU arms the identical bytes). Returns (bases uint8 cuda tensor, seg_off uint64 numpy, edge_seg_off uint32 numpy)."""
    import synth
    DEPTH = globals()["DEPTH"] if DEPTH is None else DEPTH
    GAP_LEN = globals()["GAP_LEN"] if GAP_LEN is None else GAP_LEN
    return synth.hashed_batch(n_edges, seed, depth=DEPTH, length=GAP_LEN, err=ERR, device=device, chunk=chunk)


SHARD_SEED = 1000          # rank r's cfg3 shard is hashed_batch(seed SHARD_SEED + r); the reference arm samples rank 0's


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
    # the first n edges of rank 0's shard of the GPU arm: same generator, same seed, same bytes
    bases, seg_off, eso = synth.hashed_batch(n, SHARD_SEED, depth=DEPTH, length=GAP_LEN, err=ERR)
    for _ in range(args.warmup):
        run_oracle_sample(bases, seg_off, eso, n, cores)
    t = time.perf_counter()
    nb = 0
    for _ in range(args.steps):
        _, _, _, _, b = run_oracle_sample(bases, seg_off, eso, n, cores)
        nb += b
    dt = time.perf_counter() - t
    v = nb / dt / 1e6
    sample = (f"first {n} of rank 0's {N_EDGES} cfg3 edges per step ({DEPTH} x {GAP_LEN} bp; the same bytes the GPU arm processes), "
              "restated SPOA 1.1.3 int16 SSE4.1, one edge per queue grab")
    wp = None
    if not args.no_whole_path:
        wp = reference_whole_path(cores)
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
        "data": "synthetic", "config": {"workload": f"cfg3 batched-POA stress: {N_EDGES} edges x {DEPTH} x {GAP_LEN} bp (bounded sample per step)",
                                        "scores": list(SCORES)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "whole_path": wp,
    }))


CFG2 = dict(genome=10_000_000, reads=50_000, read_len=8000, seed=1)      # BASELINE config 2 (SURVEY 8d generator)

# Strong-scaling edge population: ONE fixed set of backbone edges, modelled on the spread the reference's own run of config 2 shows
# (30 supporting reads per edge on average, p95 82; gaps median 470 bp, p95 2.5 kb, max 7.5 kb), x STRONG_SCALE towards
# config 4 (D. melanogaster = config 2 x 14). (supporting reads, share) x (gap length, share):
STRONG_DEPTHS = ((8, 0.15), (18, 0.25), (28, 0.30), (40, 0.18), (60, 0.09), (82, 0.03))
STRONG_GAPS = ((120, 0.32), (400, 0.32), (900, 0.22), (2200, 0.10), (4500, 0.033), (7500, 0.007))
STRONG_MAX_CELLS = 1.2e9     # the reference's config 2 run has no edge above 8e8 DP cells: long gaps come with few supporting reads
STRONG_EDGES = 6033 * 2


def strong_population(n_edges):
    """[(depth, gap, count, first global edge id)] and the estimated time (sharding.alignment_cost) of one edge of each class."""
    from haslr_b200 import sharding
    combos = []
    for d, wd in STRONG_DEPTHS:
        for g, wg in STRONG_GAPS:
            v, cells, work = float(g), 0.0, 0.0
            for _k in range(1, d):
                cells += v * g; work += float(sharding.alignment_cost(v, g)); v += 0.07 * g
            if cells <= STRONG_MAX_CELLS:
                combos.append((d, g, wd * wg, work))
    tot = sum(w for _, _, w, _ in combos)
    pop, cost, first = [], [], 0
    for d, g, w, work in combos:
        c = int(round(n_edges * w / tot))
        if c:
            pop.append((d, g, c, first)); cost.append(work); first += c
    return pop, cost


def strong_scaling_leg(ctx, dist, rank, world, dev, args):
    """One fixed edge set dealt to the ranks by estimated cost (sharding.shard_edges, longest processing time first), every rank
    runs hgpu_poa_batch_dev on its shard, one all-gather of the consensus; time = max over ranks. The analogue of the
    reference's only parallelism: T threads pulling the edges of ONE dataset (Assemble.cpp:386-434,562-605)."""
    import torch
    import synth
    from haslr_b200 import sharding
    pop, ccost = strong_population(STRONG_EDGES)
    n_total = sum(c for _, _, c, _ in pop)
    cost = np.concatenate([np.full(c, ccost[i]) for i, (_, _, c, _) in enumerate(pop)])
    mine = sharding.shard_edges(cost, world)[rank]
    parts, so_parts, eo_parts = [], [], []
    for d, g, c, first in pop:
        ids = mine[(mine >= first) & (mine < first + c)]
        if len(ids) == 0:
            continue
        b, so, eo = synth.hashed_batch(0, 4242 + d * 100000 + g, depth=d, length=g, err=ERR, device=dev, eids=ids - first, chunk=max(1, 60_000_000 // (d * g * 8)))
        parts.append(b); so_parts.append(np.diff(so.astype(np.int64))); eo_parts.append(np.full(len(ids), d, dtype=np.int64))
    d_bases = torch.cat(parts)
    seg_off = np.concatenate(([0], np.cumsum(np.concatenate(so_parts)))).astype(np.uint64)
    eso = np.concatenate(([0], np.cumsum(np.concatenate(eo_parts)))).astype(np.uint32)
    nb = int(seg_off[-1])
    d_out = torch.empty(nb // 4 + 65536, dtype=torch.uint8, device=dev)
    passes = 2

    def one():
        off, status = ctx.poa_batch_dev(d_bases.data_ptr(), seg_off, eso, d_out.data_ptr(), d_out.numel(), *SCORES)
        if world > 1:
            sharding.all_gather_consensus(dist, d_out, off, dev, to_host=False)
        return off, status
    off, status = one()                                  # warm-up: sizes the arenas of this shard
    assert (status == 0).all(), f"strong-scaling shard: status {np.unique(status)}"
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = 0.0
    e0.record()
    for _ in range(passes):
        one()
        kms += ctx.poa_stats()["ms_dp"]
    e1.record(); torch.cuda.synchronize()
    st = ctx.poa_stats()
    t = torch.tensor([e0.elapsed_time(e1) / passes, kms / passes, float(st["cells"]), float(nb), float(len(eso) - 1)], dtype=torch.float64, device=dev)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    rows = [x.cpu().tolist() for x in allt]
    ms = max(r[0] for r in rows)
    km = [r[1] for r in rows]
    tot_b, tot_c = sum(r[3] for r in rows), sum(r[2] for r in rows)
    return {"scaling": "strong", "workload": f"{n_total} backbone edges in {len(pop)} (supporting reads x gap) classes modelled on config 2's spread, "
                                             f"{tot_b / 1e6:.0f} Mbases, {tot_c:.3e} DP cells; the SAME set at every N, dealt by estimated cost (LPT)",
            "value": tot_b / 1e6 / (ms / 1e3), "unit": UNIT, "ms_per_pass": ms, "gcups": tot_c / (ms / 1e3) / 1e9, "n_gpus": world,
            "per_rank_ms": [r[0] for r in rows], "per_rank_kernel_ms": km, "per_rank_edges": [int(r[4]) for r in rows],
            "per_rank_cells": [r[2] for r in rows], "kernel_imbalance_max_over_mean": max(km) / (sum(km) / len(km)) if sum(km) > 0 else None,
            "limiter": "the critical path of the heaviest edges: the alignments of one edge are a serial chain (fill ~0.7 us per graph row, then "
                       "traceback, graph update and topological sort on one warp), ~12 ms per alignment of a 40-read x 5-kb edge = ~0.5 s for "
                       "the edge, whatever the number of GPUs. k_poa_pool runs such edges first and at their own speed (longest remaining "
                       "critical path first), so a rank's time is max(its share of the work, that chain); the deal's imbalance "
                       "(kernel_imbalance_max_over_mean) and the all-gather (< 1 % of a pass) do not matter before that"}





def cfg2_dataset():
    """BASELINE config 2 written by the seeded generator binary (tools/gen_synth.cpp) into a scratch directory."""
    import tempfile
    gen = next((p for p in (os.path.join(ROOT, "bin", "gen_synth"), os.path.join(ROOT, "oracle", "_ref", "gen_synth")) if os.path.exists(p)), None)
    if gen is None:
        return None
    d = tempfile.mkdtemp(prefix="haslr_cfg2_")
    subprocess.run([gen, d, str(CFG2["genome"]), str(CFG2["reads"]), str(CFG2["read_len"]), str(CFG2["seed"])], check=True, stdout=subprocess.DEVNULL)
    return d


def reference_whole_path(cores, d=None):
    """The reference binary (oracle/_ref/haslr_assemble_ref: the unmodified reference sources + the restated SPOA) on config 2, all
    host threads; stage times from its own log. Mbases = the long-read bases it fed to POA (the `>` records of log_consensus.txt)."""
    import re
    import shutil
    ref = os.path.join(ROOT, "oracle", "_ref", "haslr_assemble_ref")
    if not os.path.exists(ref):
        return {"unavailable": "oracle/_ref/haslr_assemble_ref not built"}
    own = d is None
    d = d or cfg2_dataset()
    if d is None:
        return {"unavailable": "no generator binary"}
    out = os.path.join(d, "ref")
    t = time.perf_counter()
    r = subprocess.run([ref, "-t", str(cores), "-c", "contigs.fa", "-l", "reads.fa", "-m", "map.paf", "--aln-block", "500", "--aln-sim", "0.85",
                        "--edge-sup", "3", "-d", "ref"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t
    laps = [float(x) for x in re.findall(r"elapsed time [0-9.]+ CPU seconds \(([0-9.]+) real seconds\)", r.stderr)]
    bases = 0
    with open(os.path.join(out, "log_consensus.txt"), "rb") as f:
        for ln in f:
            if ln[:1] == b">" and ln[:2] != b">C":
                bases += int(ln.split()[-1])
    # laps: contigs, kmer freq, long reads, alignment loaded, overlaps fixed, compact reads, graph, weak, tips, 3 x bubbles, coordinates, consensus, ...
    res = {"unavailable": "unexpected log"}
    if r.returncode == 0 and len(laps) >= 14:
        t_text, t_hits, t_cons = laps[2], laps[3], laps[13]
        res = {"value": bases / 1e6 / (t_cons - t_text), "unit": UNIT, "cores": cores, "kind": "reference",
               "Mbases": bases / 1e6, "s_from_paf_text": t_cons - t_text, "s_from_hits": t_cons - t_hits, "s_consensus_stage": laps[13] - laps[12],
               "s_binary_wall": wall, "sample": "one run of oracle/_ref/haslr_assemble_ref on the whole config 2 dataset (restated SPOA linked in)"}
    if own:
        shutil.rmtree(d, ignore_errors=True)
    return res


def whole_path_leg(ctx, args, peak):
    """BASELINE config 2 through libhaslr_path.so: PAF text + sequences in host memory -> consensus of every edge in host memory
    (K0 tokenise, K1 compact reads, K2 edge table, host cleaning, K4 coordinates, segment gather, K3 POA), per-stage wall time and
    per-kernel CUDA-event time against the algorithmic bytes of DESIGN.md."""
    import ctypes as C
    import shutil
    lib_p = os.environ.get("HASLR_PATH_LIB") or os.path.join(ROOT, "haslr_b200", "libhaslr_path.so")   # override: developer A/B builds
    d = cfg2_dataset()
    if d is None or not os.path.exists(lib_p):
        return {"unavailable": "generator binary or libhaslr_path.so not built"}

    class Res(C.Structure):
        _fields_ = [("n_rows", C.c_uint64), ("n_edges", C.c_uint64), ("poa_bases", C.c_uint64), ("cons_bytes", C.c_uint64), ("cons_crc", C.c_uint32),
                    ("pad", C.c_uint32)] + [(n, C.c_double) for n in ("s_tokenize", "s_k1", "s_k2", "s_clean", "s_coords", "s_poa", "s_total")]
    L = C.CDLL(lib_p)
    L.haslr_path_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.haslr_path_run.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(Res)]
    L.haslr_path_close.argtypes = [C.c_void_p]
    L.haslr_path_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4
    h = C.c_void_p()
    j = lambda n: os.path.join(d, n).encode()
    assert L.haslr_path_open(j("contigs.fa"), j("reads.fa"), j("map.paf"), C.byref(h)) == 0
    sz = [C.c_uint64() for _ in range(4)]
    L.haslr_path_sizes(h, *[C.byref(x) for x in sz])
    ctxs = (C.c_void_p * 1)(ctx.h)
    ctx.set_timing(True)
    threads = os.cpu_count() or 1
    res = Res()
    runs = []
    for it in range(1 + max(1, min(args.steps, 3))):               # one warm-up pass (sizes the device buffers), then the timed passes
        s0 = ctx.stage_stats()
        l0 = ctx.launch_count()
        t = time.perf_counter()
        rc = L.haslr_path_run(h, ctxs, 1, threads, None, C.byref(res))
        dt = time.perf_counter() - t
        assert rc == 0, ctx.L.hgpu_last_error(ctx.h).decode()
        s1 = ctx.stage_stats()
        runs.append((dt, {k: getattr(res, k) for k, _ in Res._fields_}, s1, ctx.poa_stats(), s1["h2d_bytes"] - s0["h2d_bytes"],
                     s1["d2h_bytes"] - s0["d2h_bytes"], ctx.launch_count() - l0))
    crcs = {r[1]["cons_crc"] for r in runs}
    timed = runs[1:]
    dt = sum(r[0] for r in timed) / len(timed)
    _, rr, st, ps, h2d, d2h, launches = timed[-1]
    mb = rr["poa_bases"] / 1e6
    poa_h2d = rr["poa_bases"] + 8 * (ps["alignments"] + rr["n_edges"]) + 4 * rr["n_edges"]     # hgpu_poa_batch copies its own inputs (not counted by the stage counters)
    kern = []

    def k(name, ms, n_launch, alg_bytes, units):
        kern.append({"kernel": name, "ms": ms, "launches": n_launch, "algorithmic_bytes": alg_bytes, "units": units,
                     "achieved": (alg_bytes / (ms / 1e3) / 1e9) if ms > 0 else None, "unit": "GB/s",
                     "frac": (alg_bytes / (ms / 1e3) / 1e9 / peak) if ms > 0 else None})
    k("k0_* PAF tokeniser", st["ms_k0"], st["launches_k0"], 4 * st["k0_text_bytes"] + 42 * st["k0_rows"] + 4 * st["k0_ops"],
      f"{st['k0_text_bytes']} text bytes x 4 passes + 42 B x {st['k0_rows']} rows + 4 B x {st['k0_ops']} CIGAR runs")
    k("k1_filter + k1_cigar_totals + k1_tail + k1_pack", st["ms_k1"], st["launches_k1"], 44 * st["k1_hits"] + 8 * int(sz[0].value),
      f"44 B x {st['k1_hits']} hits + 8 B x {int(sz[0].value)} contigs (SURVEY 8d)")
    k("k2_* edge table", st["ms_k2"], st["launches_k2"], 52 * st["k2_pairs"], f"52 B x {st['k2_pairs']} adjacent pairs (SURVEY 8d)")
    k("k4_edge_coords", st["ms_k4"], st["launches_k4"], 124 * st["k4_supports"] + 4 * st["k4_runs"],
      f"(24 out + 2 x 44 element + 12 in) B x {st['k4_supports']} supports + 4 B x {st['k4_runs']} CIGAR runs in the windows walked")
    k("k_poa_* (deep / team / shallow classes side by side)", ps["ms_dp"], ps["dp_launches"], 4 * ps["cells"] + ps["bases_in"] + ps["bases_out"],
      f"4 B x {ps['cells']} DP cells + bases in + consensus out (SURVEY 8d)")
    kern[-1]["gcups"] = ps["cells"] / (ps["ms_dp"] / 1e3) / 1e9 if ps["ms_dp"] > 0 else None
    out = {
        "workload": ("BASELINE config 2" if CFG2["genome"] == 10_000_000 else f"BASELINE config 2's generator x {CFG2['genome'] // 10_000_000} "
                     "(x 14 = config 4, D. melanogaster scale)") +
                    f": synthetic {CFG2['genome'] // 1_000_000} Mb genome, {int(sz[0].value)} SRC contigs, {int(sz[1].value)} long reads "
                    f"({int(sz[2].value) / 1e6:.0f} Mbases), {int(sz[3].value) / 1e6:.0f} MB of PAF text ({rr['n_rows']} rows), seed {CFG2['seed']}",
        "boundary": "haslr_path_run (libhaslr_path.so above the C ABI): inputs in HOST memory -> consensus strings in HOST memory",
        "value": mb / dt, "unit": UNIT, "Mbases": mb, "s_per_pass": dt, "passes_timed": len(timed), "edges": rr["n_edges"],
        "from_hits": {"value": mb / (dt - sum(r[1]["s_tokenize"] for r in timed) / len(timed)), "unit": UNIT,
                      "note": "the same passes without the PAF tokenising stage (SURVEY 8d starts its clock with the hits parsed)"},
        "stage_wall_s": {n[2:]: sum(r[1][n] for r in timed) / len(timed) for n in ("s_tokenize", "s_k1", "s_k2", "s_clean", "s_coords", "s_poa")},
        "h2d_bytes_per_pass": h2d + poa_h2d, "d2h_bytes_per_pass": d2h + rr["cons_bytes"] + 12 * rr["n_edges"],
        "cg_ops_uploads_per_pass": 0, "paf_text_uploads_per_pass": 1, "gpu_launches_per_pass": launches,
        "kernels": kern, "deterministic": len(crcs) == 1, "consensus_crc32": "%08x" % rr["cons_crc"],
        "check": "tests/test_drop_in_gpu.py::test_baseline_config2_full_size compares every output file of this dataset with the reference binary's",
    }
    if args.whole_path_ref:
        out["cpu_reference"] = reference_whole_path(threads, d)
    ctx.set_timing(False); ctx.poa_set_timing(True)
    L.haslr_path_close(h)
    shutil.rmtree(d, ignore_errors=True)
    return out


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
    ap.add_argument("--no-whole-path", action="store_true", help="skip the whole-path leg (BASELINE config 2: PAF text -> consensus)")
    ap.add_argument("--whole-path-ref", action="store_true", help="also time the reference binary on the whole-path dataset inside the native arm")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (one fixed edge set dealt to the N ranks)")
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
    d_bases, seg_off, eso = gen_cfg3_torch(n_edges, SHARD_SEED + rank, dev)
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

    # ---- strong scaling: one fixed edge set dealt over the ranks (every rank takes part; rank 0 reports)
    strong = None
    if not args.no_strong:
        strong = strong_scaling_leg(ctx, dist if world > 1 else None, rank, world, dev, args)

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
                "algorithmic_bytes_per_launch": alg_bytes / max(1, args.steps * klaunch),
                # frac can exceed 1: SURVEY 8d counts an int32 cell (4 B), the kernel stores int16 cells (2 B) and reads a fraction back,
                # so the DRAM pipe itself moves `traffic` bytes per launch, not the algorithmic figure
                "dram_frac": (traffic / (kernel_ms / 1e3 / max(1, args.steps * klaunch)) / 1e9 / peak) if traffic else None,
                "note": "achieved = SURVEY 8d algorithmic bytes (4 B per DP cell) / kernel time; dram_frac = ncu DRAM bytes of the same cells / "
                        "kernel time / peak. The kernel is issue-bound (ncu: ALU pipe 58 %, DRAM 49 %), see DESIGN.md 2d"}

    # ---- second shape, reported beside the headline: deep edges as a real 25x dataset produces them (BASELINE config 2:
    #      median 28 supporting reads over a 2.5 kb gap, graphs of ~10^4 nodes, most cells outside the plain int16 range)
    deep = None
    if world == 1 and not args.no_deep:
        deep = {}
        for n_deep_edges, key in ((DEEP_EDGES, None), (4 * DEEP_EDGES, "saturated")):
            dd, dso, deo = gen_cfg3_torch(n_deep_edges, 77, dev, chunk=64, DEPTH=DEEP_DEPTH, GAP_LEN=DEEP_GAP)
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
            if not args.no_cpu and key is None:
                rc, roff, ccells, cdt, cnb = run_oracle_sample(dd.cpu().numpy(), dso, deo, 4, os.cpu_count() or 1)
                assert np.array_equal(doff[:5], roff) and dout[: int(doff[4])].cpu().numpy().tobytes() == rc.tobytes(), "deep edges: consensus differs from the oracle"
                dchk = "first 4 edges bit-exact vs oracle"
                # the checker's own speed on those 4 edges (one edge per thread, so 4 host threads busy): context, not a baseline run
                dcpu = {"edges": 4, "threads_busy": 4, "value": cnb / cdt / 1e6, "unit": UNIT, "gcups": ccells / cdt / 1e9}
            leg = {"workload": f"{n_deep_edges} edges x {DEEP_DEPTH} supporting reads x {DEEP_GAP} bp gap (BASELINE config 2 edge shape, "
                               f"{n_deep_edges // 148} edges per SM)",
                   "value": dn / (dms / 1e3) / 1e6, "unit": UNIT, "ms_per_step": dms, "gcups": dst["cells"] / (dms / 1e3) / 1e9,
                   "kernel_ms": dst["ms_dp"], "alignments": dst["alignments"], "alignments_rel16": dst["alignments_rel16"],
                   "alignments_i32": dst["alignments_i32"], "kernels": "k_poa_pool", "check": dchk, "cpu_oracle_on_check": dcpu}
            if key is None:
                deep = leg
            else:
                deep[key] = leg
            del dd, dout
            torch.cuda.empty_cache()

    # ---- the whole path (BASELINE config 2) through the path library
    whole = None
    if world == 1 and not args.no_whole_path:
        whole = whole_path_leg(ctx, args, peak)

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
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "deep_edges": deep, "whole_path": whole, "strong_scaling": strong,
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
