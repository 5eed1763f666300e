"""Seeded random inputs of the edge-coordinate stage that stress what the golden dataset barely touches: many supports
per edge (several bitmask words), ties between begin and end positions, equal-depth optima (the >= / > difference of the
two sweeps), walks that are refused (-1), rows without CIGAR, both strands of both anchors."""
import numpy as np

import oracle_ffi


def random_case(seed, n_edges=200, max_supp=90, n_reads=400):
    rng = np.random.default_rng(seed)
    hits = dict(is_rev=[], cg_off=[0], cg_ops=[])
    elems, cl_off, read_len = [], [0], []
    per_read = []
    for r in range(n_reads):
        ne = int(rng.integers(2, 7))
        rl = int(rng.integers(3000, 12000))
        read_len.append(rl)
        per_read.append(ne)
        q = int(rng.integers(0, 200))
        for _ in range(ne):
            h = len(hits["is_rev"])
            hits["is_rev"].append(int(rng.integers(0, 2)))
            # run-length CIGAR: M/I/D runs; sometimes none at all (row without cg:Z:)
            runs = []
            if rng.random() > 0.05:
                for k in range(int(rng.integers(1, 40))):
                    op = 0 if k % 2 == 0 else int(rng.integers(1, 3))
                    runs.append((int(rng.integers(1, 60 if op == 0 else 6)) << 2) | op)
            hits["cg_ops"] += runs
            hits["cg_off"].append(len(hits["cg_ops"]))
            m = sum(x >> 2 for x in runs if x & 3 == 0); ins = sum(x >> 2 for x in runs if x & 3 == 1); d = sum(x >> 2 for x in runs if x & 3 == 2)
            t_start = int(rng.integers(0, 40)) if rng.random() < 0.7 else int(rng.integers(0, 400))
            t_end = t_start + max(1, m + d)
            lo, hi = 0, max(0, len(runs) - 1)
            lo_len = (runs[lo] >> 2) if runs else 0
            hi_len = (runs[hi] >> 2) if runs else 0
            if runs and rng.random() < 0.3 and lo_len > 1:      # a window trimmed by the overlap fix
                lo_len = int(rng.integers(1, lo_len))
            q_end = q + max(1, m + ins)
            elems.append((h, q, min(q_end, rl), t_start, t_end, m, m + ins + d, lo, lo_len if lo != hi else min(lo_len, hi_len) if runs else 0, hi, hi_len))
            q = q_end + int(rng.integers(0, 50))
        cl_off.append(len(elems))
    elems = np.array(elems, dtype=oracle_ffi.CL_ELEM)
    supp, supp_off, edge_rev = [], [0], []
    for e in range(n_edges):
        n = int(rng.integers(1, max_supp))
        # quantise positions so begins and ends collide across supports
        for _ in range(n):
            r = int(rng.integers(0, n_reads))
            a = int(rng.integers(0, per_read[r] - 1))
            if rng.random() < 0.5:
                supp.append((r | (int(rng.integers(0, 2)) << 31), a, a + 1))
            else:
                supp.append((r, a + 1, a))
        supp_off.append(len(supp))
        edge_rev.append(int(rng.integers(0, 4)))
    if hits["cg_ops"] == []:
        hits["cg_ops"] = [0]
    h = dict(is_rev=np.array(hits["is_rev"], dtype=np.uint8), cg_off=np.array(hits["cg_off"], dtype=np.uint32),
             cg_ops=np.array(hits["cg_ops"], dtype=np.uint32))
    # ties: snap a third of the element coordinates onto a coarse grid
    snap = rng.random(len(elems)) < 0.35
    elems["t_start"][snap] = (elems["t_start"][snap] // 16) * 16
    elems["t_end"][snap] = np.maximum(elems["t_start"][snap] + 1, (elems["t_end"][snap] // 16) * 16)
    # degenerate elements (empty or inverted target interval): the sweep then erases a support before it inserts it, and walks are refused
    deg = rng.random(len(elems)) < 0.02
    elems["t_end"][deg] = np.maximum(elems["t_start"][deg].astype(np.int64) - rng.integers(0, 3, int(deg.sum())), 0).astype(np.uint32)
    return dict(edge_rev=np.array(edge_rev, dtype=np.uint8), supp_off=np.array(supp_off, dtype=np.uint32),
                supp=np.array(supp, dtype=oracle_ffi.EDGE_SUPP), elems=elems, cl_off=np.array(cl_off, dtype=np.uint32),
                read_len=np.array(read_len, dtype=np.uint32), hits=h)
