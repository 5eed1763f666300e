#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): headline metrics + warp-stall samples / instructions per source phase.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.json]"""
import collections
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.sum']
head = {h: (vals[i], units[i]) for i, h in enumerate(hdr) if h in keep}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
lines = collections.OrderedDict()
ts = ti = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name') or r[0] == '' or r[2] != '-':
        continue
    try:
        samp, inst = int(r[4]), int(r[7])
    except ValueError:
        continue
    a = lines.setdefault((cur, int(r[0])), [0, 0, r[1][:110]])
    a[0] += samp; a[1] += inst; ts += samp; ti += inst
print(json.dumps(head, indent=1))
print("total samples", ts, "instructions", ti)
print("--- top lines by stall samples")
for (f, l), (s, i, t) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{f}:{l:4d} samp {100*s/ts:5.1f}% inst {100*i/ti:5.1f}%  {t}")
if len(sys.argv) > 2 and sys.argv[2] != '-':
    json.dump({"headline": head, "total_samples": ts, "total_inst": ti,
               "top_lines": [{"file": f, "line": l, "samples_pct": 100*s/ts, "inst_pct": 100*i/ti, "src": t}
                             for (f, l), (s, i, t) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:60]]}, open(sys.argv[2], "w"), indent=1)

# ---- per-function / per-phase attribution (function = nearest preceding definition in the same file)
import os
import re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
defs = {}
SUBPHASE = [("--- sequence profile of this stripe", "fill.profile"), ("--- row 0: Hhat = 0", "fill.row0"),
            ("batched per-rank records", "fill.batch_hdr"), ("const uint32_t i = r0 + q + 1;", "fill.row_hdr"),
            ("if ((m0 & META_FAST) != 0) {", "fill.fast_row"), ("const uint32_t npc = (m0 >> 3) & 3u, d0 = m0 >> META_D0_SHIFT;", "fill.slow_row"),
            ("// horizontal gaps = prefix maximum", "fill.scan"), ("// stream the row out", "fill.store"),
            ("// end cell: best Hhat", "tb.endcell"), ("// tile of the stored matrix", "tb.tile_load"),
            ("// (1) a run of diagonal moves", "tb.diag_run"), ("// (2) one generic step", "tb.generic"),
            ("// SPOA's DFS from root", "topo.dfs"), ("// park row i-1: later rows read it", "rel.row_head"), ("// the first predecessor initialises the row", "rel.first_pred"),
            ("// the others (never the previous rank", "rel.more_preds"), ("// generic walk: predecessors from the CSR", "rel.generic_row"),
            ("// horizontal gaps = prefix maximum in hat space (see row16)", "rel.scan_store"), ("// --- row 0: Hhat = 0 everywhere, base 0", "rel.stripe_setup"), ("// (A)+(B): node of every position", "add.AB"), ("// (B2): initialise new nodes", "add.B2"),
            ("// (C): edges between consecutive", "add.C"), ("// heaviest-bundle consensus; node ids", "edge.consensus"), ("// publish", "edge.publish")]
for fn in ("poa_device.cuh", "poa_graph.cuh", "poa_fill_rel.cuh", "poa_pool.cuh"):
    marks = []
    for n, text in enumerate(open(os.path.join(ROOT, "haslr_b200", "csrc", fn)), 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|HGPU_HD|__host__ __device__)[^;]*?\b([a-zA-Z_0-9]+)\s*\(", text)
        if m and not text.strip().endswith(";"):
            marks.append((n, m.group(1)))
        for pat, nm in SUBPHASE:
            if pat in text:
                marks.append((n, nm))
    defs[fn] = marks
phase = collections.Counter(); phase_i = collections.Counter()
for (f, l), (s, i, t) in lines.items():
    name = f or "?"
    for n, nm in defs.get(f, []):
        if n <= l:
            name = nm
    phase[name] += s; phase_i[name] += i
print("--- by function/phase (samples %, instructions %)")
for k, v in phase.most_common(40):
    print(f"{k:28s} samp {100*v/ts:5.1f}%  inst {100*phase_i[k]/ti:5.1f}%")
