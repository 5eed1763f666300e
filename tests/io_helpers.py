"""Test-side readers/writers for the file formats around the hot path.

Plain Python/numpy, used to feed the oracle and the C-ABI with identical arrays and to render
their outputs in the reference's text formats for byte comparison with tests/golden/.
"""
import re

import numpy as np

_CG = re.compile(rb"(\d+)([A-Za-z=])")


def load_contigs(path):
    """contigs.fa -> (len uint32[], mean_kmer float64[], kmer_count uint32[], seqs list[bytes]). Contig.cpp:43-107."""
    lens, km, kc, seqs = [], [], [], []
    with open(path, "rb") as f:
        name = None
        buf = []
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    seqs.append(b"".join(buf))
                name = line
                kc.append(int(re.search(rb"KC:i:(\d+)", line).group(1)))
                km.append(float(re.search(rb"km:f:([0-9.eE+-]+)", line).group(1)))
                buf = []
            else:
                buf.append(line.strip())
        if name is not None:
            seqs.append(b"".join(buf))
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    return lens, np.array(km, dtype=np.float64), np.array(kc, dtype=np.uint32), seqs


def calc_uniq_freq(lens, km):
    """Contig.cpp:162-174: mean km of the 20 longest contigs (pairs sorted descending by (len, km))."""
    order = sorted(zip(lens.tolist(), km.tolist()), reverse=True)[:20]
    freq = 0.0
    for _, k in order:
        freq += k
    return freq / len(order)


def load_fasta(path):
    seqs = []
    with open(path, "rb") as f:
        buf = None
        for line in f:
            if line.startswith(b">"):
                if buf is not None:
                    seqs.append(b"".join(buf))
                buf = []
            else:
                buf.append(line.strip())
        if buf is not None:
            seqs.append(b"".join(buf))
    return seqs


def parse_paf(path, n_reads):
    """map.paf -> (hits dict of SoA numpy arrays, read_off uint32[n_reads+1]). Columns as Longread.cpp:275-289.

    Run-length CIGAR ops are packed (len << 2) | op with op 0 = M, 1 = I, 2 = anything else.
    Hits must be grouped by read in ascending read id (the reference's silent assumption).
    """
    cols = {k: [] for k in ("q_id", "q_len", "q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block")}
    is_rev, mapq, cg_off, cg_ops = [], [], [0], []
    opmap = {b"M": 0, b"I": 1}
    with open(path, "rb") as f:
        for line in f:
            fl = line.rstrip(b"\n").split(b"\t")
            cols["q_id"].append(int(fl[0])); cols["q_len"].append(int(fl[1]))
            cols["q_start"].append(int(fl[2])); cols["q_end"].append(int(fl[3]))
            is_rev.append(1 if fl[4][:1] == b"-" else 0)
            cols["t_id"].append(int(fl[5])); cols["t_len"].append(int(fl[6]))
            cols["t_start"].append(int(fl[7])); cols["t_end"].append(int(fl[8]))
            cols["n_match"].append(int(fl[9])); cols["n_block"].append(int(fl[10]))
            mapq.append(int(fl[11]) & 0xFF)
            for x in fl[12:]:
                if x.startswith(b"cg:Z:"):
                    for n, op in _CG.findall(x[5:]):
                        cg_ops.append((int(n) << 2) | opmap.get(op, 2))
                    break
            cg_off.append(len(cg_ops))
    h = {k: np.array(v, dtype=np.uint32) for k, v in cols.items()}
    h["is_rev"] = np.array(is_rev, dtype=np.uint8)
    h["mapq"] = np.array(mapq, dtype=np.uint8)
    h["cg_off"] = np.array(cg_off, dtype=np.uint32)
    h["cg_ops"] = np.array(cg_ops if cg_ops else [0], dtype=np.uint32)
    q = h["q_id"]
    assert np.all(np.diff(q.astype(np.int64)) >= 0), "PAF must list reads in ascending id"
    read_off = np.searchsorted(q, np.arange(n_reads + 1), side="left").astype(np.uint32)
    return h, read_off


def format_compact(elems, read_off, hits):
    """compact_uniq.txt exactly as print_compact_longreads (Longread.cpp:675-693)."""
    out = []
    tid = hits["t_id"]; rev = hits["is_rev"]
    for r in range(len(read_off) - 1):
        parts = [">%d\t" % r]
        for e in elems[read_off[r]:read_off[r + 1]]:
            parts.append("%d-%d:%d:%s:%d-%d\t" % (e["q_start"], e["q_end"], tid[e["hit"]], "-" if rev[e["hit"]] else "+",
                                                  e["t_start"], e["t_end"]))
        parts.append("\n")
        out.append("".join(parts))
    return "".join(out)


def format_gfa_links(keys):
    """The L lines of bbg_print_graph_gfa (Backbone_graph.cpp:576-585) for directed entries sorted by key64."""
    out = []
    for k in keys.tolist():
        frm, to = k >> 32, k & 0xFFFFFFFF
        out.append("L\t%d\t%s\t%d\t%s\t0M\n" % (frm >> 1, "+-"[frm & 1], to >> 1, "+-"[to & 1]))
    return "".join(out)


def gfa_links_of_file(path):
    with open(path) as f:
        return "".join(l for l in f if l.startswith("L\t"))


def reduce_coordinate_log(path):
    """log_coordinate.txt of the reference (asm_calc_single_edge_coordinates, Assemble.cpp:157-363) reduced to:
    per edge  E node1 rev1 node2 rev2 n_supp int1_lo int1_hi int2_lo int2_hi c1 c2 n_best,
    per support  D head(t_start t_end strand) tail(t_start t_end strand),
    per best support  S lr len strand spos epos | S lr len strand X."""
    sg = {"+": 0, "-": 1}
    out = []
    cur = None

    def flush():
        if cur:
            out.append("E %s %s\n" % (" ".join(map(str, cur["e"])), " ".join(map(str, cur["v"]))))
            out.extend(cur["lines"])
    with open(path) as f:
        for line in f:
            t = line.split()
            if line.startswith("edge "):
                flush()
                a, b = t[1].split(":"), t[3].split(":")
                cur = {"e": [int(a[0]), sg[a[1]], int(b[0]), sg[b[1]]], "v": [], "lines": []}
            elif line.startswith("\tedge_supp size:"):
                cur["v"].append(int(line.split(":")[1]))
            elif line.startswith("\tsupp_detail"):
                cur["lines"].append("D %s %s %d %s %s %d\n" % (t[2], t[3], sg[t[4]], t[6], t[7], sg[t[8]]))
            elif "@@@" in line:
                cur["v"] += [int(t[-2]), int(t[-1])]
            elif line.startswith("coordinates"):
                cur["v"] += [int(t[2]), int(t[4])]
            elif line.startswith("supproting_lr"):
                cur["v"].append(int(t[1]))
            elif "+++" in line:
                cur["pending"] = "S %s %s %d" % (t[1].split(":")[1], t[2].split(":")[1], sg[t[3].split(":")[1]])
            elif "[coordinate]" in line:
                if "could not" in line:
                    cur["lines"].append(cur["pending"] + " X\n")
                else:
                    cur["lines"].append("%s %s %s\n" % (cur["pending"], t[2].split(":")[1], t[3].split(":")[1]))
    flush()
    return "".join(out)
