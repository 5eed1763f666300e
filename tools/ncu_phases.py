#!/usr/bin/env python
"""Per-phase breakdown of an .ncu-rep of k_poa_edges (run here, no GPU): warp-stall samples by reason, instruction share and
average active threads, attributed to source regions of poa_device.cuh / poa_graph.cuh found by function name or marker.
usage: tools/ncu_phases.py report.ncu-rep [out.json]"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKS = [  # (substring that starts a region, phase name), looked up per file in source order
    ("struct SlotView", "slotview(load: end cell, generic step)"), ("struct RowOps", "fill32.rowops"), ("void team_publish", "team"),
    ("bool dp_fill(", "fill32"), ("struct Fill16", "fill16.helpers"), ("void row16(", "fill16.row:predecessors"),
    ("// ---- row i-1 (registers): diagonal and vertical moves, in place", "fill16.row:from_row_i-1"),
    ("// ---- every other predecessor row streams through two registers", "fill16.row:other_predecessor_rows"),
    ("// horizontal gaps = prefix maximum in hat space. In the lane", "fill16.row:prefix_max"),
    ("// stream the row out; its last cell", "fill16.row:store"), ("bool meta_reads_two_back", "fill16.batch"),
    ("void fill16_profile(", "fill16.profile"), ("bool dp_fill16(", "fill16.setup+batch_loop"),
    ("struct TbTile", "tb.endcell"), ("// ---- tile with (ci, cj) in its corner", "tb.tile_load+prefetch"),
    ("// ---- every lane decides the move of cell", "tb.lane_decisions"), ("// ---- kr == 2: one generic step", "tb.generic"),
    ("bool dp_align(", "dp_align"),
    ("void w_init_chain(", "init_chain"), ("void w_build_meta(", "build_meta"), ("uint32_t w_add_alignment(", "add_alignment"),
    ("int w_toposort(", "topo.batch"), ("// SPOA's DFS from root", "topo.dfs"), ("nr = __shfl_sync(FULL, nr, 0);", "topo.batch"),
    ("uint32_t w_consensus_scores(", "consensus.scores"), ("uint32_t w_consensus_backtrack(", "consensus.backtrack"),
    ("int lane_id()", "kernel.main"), ("void k_poa_edges(", "kernel.main"), ("void k_poa_edges_team(", "kernel.team"),
    ("uint32_t g_branch_completion(", "graph.consensus_serial"), ("bool g_add_alignment(", "graph.serial_add"), ("bool g_toposort(", "graph.serial_topo"),
]


def regions(path):
    text = open(path).read()
    out = []
    for pat, name in MARKS:
        for m in re.finditer(re.escape(pat), text):
            out.append((text.count("\n", 0, m.start()) + 1, name))
    return sorted(out)


def main():
    rep = sys.argv[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    reg = {fn: regions(os.path.join(ROOT, "haslr_b200", "csrc", fn)) for fn in ("poa_device.cuh", "poa_graph.cuh")}
    cur, idx = None, {}
    ph = collections.defaultdict(collections.Counter)
    keys = ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_selected", "stall_not_selected", "stall_branch_resolving", "stall_no_inst", "stall_mio", "stall_lg")
    for r in csv.reader(io.StringIO(src)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if len(r) > 2 and r[0] == "Line No":
            idx = {n: i for i, n in enumerate(r)}; continue
        if len(r) < 10 or r[0] == "" or r[2] != "-":
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        name = cur or "?"
        for a, n in reg.get(cur, []):
            if a <= ln:
                name = n

        def g(k):
            try:
                return int(r[idx[k]])
            except (ValueError, KeyError):
                return 0
        c = ph[name]
        c["samp"] += g("# Samples"); c["inst"] += g("Instructions Executed"); c["tinst"] += g("Thread Instructions Executed")
        for k in keys:
            c[k] += g(k)
    ts = sum(c["samp"] for c in ph.values()); ti = sum(c["inst"] for c in ph.values())
    print("%-34s %6s %6s %5s | %6s %6s %6s %6s %6s %6s %6s %6s" % ("phase", "samp%", "inst%", "thr", "longsb", "shrtsb", "wait", "math", "sel", "notsel", "branch", "noinst"))
    rows = []
    for n, c in sorted(ph.items(), key=lambda kv: -kv[1]["samp"]):
        row = dict(phase=n, samples_pct=100 * c["samp"] / ts, inst_pct=100 * c["inst"] / ti, threads=c["tinst"] / max(1, c["inst"]),
                   **{k: 100 * c[k] / ts for k in keys})
        rows.append(row)
        if c["samp"] >= ts * 0.002:
            print("%-34s %6.1f %6.1f %5.1f | %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f" % (
                n, row["samples_pct"], row["inst_pct"], row["threads"], row["stall_long_sb"], row["stall_short_sb"], row["stall_wait"], row["stall_math"],
                row["stall_selected"], row["stall_not_selected"], row["stall_branch_resolving"], row["stall_no_inst"]))
    tot = {k: sum(r[k] for r in rows) for k in keys}
    print("total samples", ts, "instructions", ti, "| stall totals %:", {k.replace("stall_", ""): round(v, 1) for k, v in tot.items()})
    if len(sys.argv) > 2:
        json.dump({"total_samples": ts, "total_inst": ti, "stall_totals_pct": tot, "phases": rows}, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
