"""ctypes binding of the CPU oracle (oracle/_ref/liboracle.so). TEST INFRASTRUCTURE ONLY.

The product package (haslr_b200) never imports this module; it is the checker used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "liboracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "haslr_assemble_ref")
GEN_BIN = os.path.join(ORACLE_DIR, "_ref", "gen_synth")

u8p, u32p, u64p, i32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32, C.c_double))


def build(targets=("oracle", "tools")):
    """Compile the oracle (and, when /root/reference is present, the reference binary)."""
    t = list(targets)
    if os.path.isdir("/root/reference/src/haslr_assemble/src") and "ref" not in t:
        t.append("ref")
    subprocess.run(["make", "-s", "-C", ORACLE_DIR] + t, check=True)


class HitsT(C.Structure):
    _fields_ = [("n_hits", C.c_uint32)] + [(n, u32p) for n in
                ("q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block")] + \
               [("is_rev", u8p), ("mapq", u8p), ("cg_off", u32p), ("cg_ops", u32p)]


class K1Params(C.Structure):
    _fields_ = [("min_aln_sim", C.c_double), ("uniq_freq", C.c_double), ("max_uniq_dev", C.c_double),
                ("min_aln_block", C.c_uint32), ("min_aln_mapq", C.c_uint32)]


CL_ELEM = np.dtype([(n, "<u4") for n in ("hit", "q_start", "q_end", "t_start", "t_end", "n_match", "n_block",
                                         "cg_lo", "cg_lo_len", "cg_hi", "cg_hi_len")])
EDGE_SUPP = np.dtype([("lr_id_strand", "<u4"), ("cmp_head", "<u4"), ("cmp_tail", "<u4")])


class DbgSizes(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_edges", C.c_uint32), ("aln_len", C.c_uint32), ("L", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(("oracle",))
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_poa_batch.restype = C.c_int
        _lib.oracle_poa_batch.argtypes = [u8p, u64p, u32p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          u8p, C.c_uint64, u64p, u64p, u32p]
        _lib.oracle_compact_lr.restype = C.c_int64
        _lib.oracle_compact_lr.argtypes = [C.POINTER(HitsT), u32p, C.c_uint32, f64p, C.POINTER(K1Params), C.c_void_p, u32p]
        _lib.oracle_backbone_edges.restype = C.c_int64
        _lib.oracle_backbone_edges.argtypes = [u32p, u8p, u32p, C.c_uint32, C.c_uint32, u64p, u32p, C.c_void_p, u8p]
        _lib.oracle_poa_debug.restype = C.c_int
        _lib.oracle_poa_debug.argtypes = [u8p, u64p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                          i32p, C.c_uint64, i32p, i32p, C.c_uint32,
                                          u32p, u8p, u32p, u32p, u32p, C.c_uint32, C.c_uint32, C.POINTER(DbgSizes)]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def hits_struct(h):
    """h: dict of contiguous numpy arrays (see tests/io_helpers.parse_paf). Returns (HitsT, keepalive)."""
    s = HitsT()
    s.n_hits = len(h["q_start"])
    for n in ("q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block", "cg_off", "cg_ops"):
        setattr(s, n, _p(h[n], u32p))
    s.is_rev = _p(h["is_rev"], u8p)
    s.mapq = _p(h["mapq"], u8p)
    return s


def poa_batch(bases, seg_off, edge_seg_off, match=5, mismatch=-4, gap=-8, simd=True, threads=1):
    """Returns (cons bytes array, cons_off uint64[n_edges+1], cells, nodes uint32[n_edges])."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    seg_off = np.ascontiguousarray(seg_off, dtype=np.uint64)
    edge_seg_off = np.ascontiguousarray(edge_seg_off, dtype=np.uint32)
    n_edges = len(edge_seg_off) - 1
    cap = int(seg_off[-1]) + 16
    out = np.zeros(cap, dtype=np.uint8)
    off = np.zeros(n_edges + 1, dtype=np.uint64)
    nodes = np.zeros(max(n_edges, 1), dtype=np.uint32)
    cells = C.c_uint64(0)
    rc = lib().oracle_poa_batch(_p(bases, u8p), _p(seg_off, u64p), _p(edge_seg_off, u32p), n_edges, match, mismatch, gap,
                                int(simd), threads, _p(out, u8p), cap, _p(off, u64p), C.byref(cells), _p(nodes, u32p))
    if rc != 0:
        raise RuntimeError("oracle_poa_batch failed")
    return out[: int(off[-1])], off, cells.value, nodes[:n_edges]


def poa_debug(bases, seg_off, n_prior, match=5, mismatch=-4, gap=-8, want_H=True):
    """Graph after n_prior segments (rank order) + H/alignment of the next segment."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    seg_off = np.ascontiguousarray(seg_off, dtype=np.uint64)
    n_segs = len(seg_off) - 1
    cap_n = int(seg_off[-1]) + 1
    lens = np.diff(seg_off.astype(np.int64))
    Lmax = int(lens.max()) if n_segs else 0
    H = np.zeros((cap_n + 1) * (Lmax + 1) if want_H else 1, dtype=np.int32)
    aln_cap = cap_n + Lmax + 2
    an = np.zeros(aln_cap, dtype=np.int32); ap = np.zeros(aln_cap, dtype=np.int32)
    r2n = np.zeros(cap_n, dtype=np.uint32); code = np.zeros(cap_n, dtype=np.uint8)
    poff = np.zeros(cap_n + 1, dtype=np.uint32); pn = np.zeros(cap_n + n_segs + 8, dtype=np.uint32); pw = np.zeros_like(pn)
    sz = DbgSizes()
    rc = lib().oracle_poa_debug(_p(bases, u8p), _p(seg_off, u64p), n_segs, n_prior, match, mismatch, gap,
                                _p(H, i32p) if want_H else None, len(H) if want_H else 0, _p(an, i32p), _p(ap, i32p), aln_cap,
                                _p(r2n, u32p), _p(code, u8p), _p(poff, u32p), _p(pn, u32p), _p(pw, u32p), cap_n, len(pn), C.byref(sz))
    if rc < 0:
        raise RuntimeError("oracle_poa_debug failed")
    V, L = sz.n_nodes, sz.L
    return dict(V=V, L=L, H=H[: (V + 1) * (L + 1)].reshape(V + 1, L + 1) if want_H else None,
                aln_node=an[: sz.aln_len].copy(), aln_pos=ap[: sz.aln_len].copy(), rank2node=r2n[:V].copy(),
                code=code[:V].copy(), pred_off=poff[: V + 1].copy(), pred_node=pn[: int(poff[V])].copy(),
                pred_weight=pw[: int(poff[V])].copy())


def compact_lr(hits, read_off, mean_kmer, uniq_freq, min_aln_block=500, min_aln_sim=0.85, min_aln_mapq=55, max_uniq_dev=0.15):
    hs = hits_struct(hits)
    read_off = np.ascontiguousarray(read_off, dtype=np.uint32)
    mean_kmer = np.ascontiguousarray(mean_kmer, dtype=np.float64)
    n_reads = len(read_off) - 1
    prm = K1Params(min_aln_sim, uniq_freq, max_uniq_dev, min_aln_block, min_aln_mapq)
    elems = np.zeros(max(hs.n_hits, 1), dtype=CL_ELEM)
    out_off = np.zeros(n_reads + 1, dtype=np.uint32)
    n = lib().oracle_compact_lr(C.byref(hs), _p(read_off, u32p), n_reads, _p(mean_kmer, f64p), C.byref(prm),
                                elems.ctypes.data, _p(out_off, u32p))
    return elems[:n].copy(), out_off


def backbone_edges(cl_tid, cl_rev, cl_read_off, min_edge_sup=3):
    cl_tid = np.ascontiguousarray(cl_tid, dtype=np.uint32)
    cl_rev = np.ascontiguousarray(cl_rev, dtype=np.uint8)
    cl_read_off = np.ascontiguousarray(cl_read_off, dtype=np.uint32)
    n_reads = len(cl_read_off) - 1
    cnt = np.diff(cl_read_off.astype(np.int64))
    n_pairs = int(np.maximum(cnt - 1, 0).sum())
    cap = max(2 * n_pairs, 1)
    key = np.zeros(cap, dtype=np.uint64); soff = np.zeros(cap + 1, dtype=np.uint32)
    supp = np.zeros(cap, dtype=EDGE_SUPP); keep = np.zeros(cap, dtype=np.uint8)
    n = lib().oracle_backbone_edges(_p(cl_tid, u32p), _p(cl_rev, u8p), _p(cl_read_off, u32p), n_reads, min_edge_sup,
                                    _p(key, u64p), _p(soff, u32p), supp.ctypes.data, _p(keep, u8p))
    return key[:n].copy(), soff[: n + 1].copy(), supp[: int(soff[n])].copy(), keep[:n].copy()


EDGE_COORD = np.dtype([(n, "<u4") for n in ("int1_lo", "int1_hi", "int2_lo", "int2_hi", "c1", "c2", "n_best", "n_cns")])
SUPP_COORD = np.dtype([("lr_start", "<i8"), ("lr_end", "<i8"), ("lr_strand", "<u4"), ("in_best", "<u4")])


def edge_coords(edge_rev, supp_off, supp, elems, cl_read_off, read_len, hits):
    """K4 oracle: per edge (rev1 | rev2 << 1) and its supports -> (EDGE_COORD[n_edges], SUPP_COORD[n_supp])."""
    edge_rev = np.ascontiguousarray(edge_rev, dtype=np.uint8)
    supp_off = np.ascontiguousarray(supp_off, dtype=np.uint32)
    supp = np.ascontiguousarray(supp, dtype=EDGE_SUPP)
    elems = np.ascontiguousarray(elems, dtype=CL_ELEM)
    cl_read_off = np.ascontiguousarray(cl_read_off, dtype=np.uint32)
    read_len = np.ascontiguousarray(read_len, dtype=np.uint32)
    n = len(edge_rev)
    oe = np.zeros(max(n, 1), dtype=EDGE_COORD); os_ = np.zeros(max(len(supp), 1), dtype=SUPP_COORD)
    L = lib()
    L.oracle_edge_coords.restype = C.c_int
    L.oracle_edge_coords.argtypes = [C.c_uint32, u8p, u32p, C.c_void_p, C.c_void_p, u32p, u32p, u8p, u32p, u32p, C.c_void_p, C.c_void_p]
    rc = L.oracle_edge_coords(n, _p(edge_rev, u8p), _p(supp_off, u32p), supp.ctypes.data, elems.ctypes.data, _p(cl_read_off, u32p),
                              _p(read_len, u32p), _p(hits["is_rev"], u8p), _p(hits["cg_off"], u32p), _p(hits["cg_ops"], u32p),
                              oe.ctypes.data, os_.ctypes.data)
    assert rc == 0
    return oe[:n], os_[: len(supp)]


PAF_COLS = ("q_id", "q_len", "q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block")


def parse_paf_with(fn, text):
    """Calls a C parser with oracle_parse_paf's signature; returns the hits dict (None if it refused the text), rc."""
    buf = np.frombuffer(bytes(text), dtype=np.uint8)
    row_cap = int((buf == 10).sum()) + 2
    op_cap = len(buf) // 2 + 2
    h = {k: np.zeros(row_cap, dtype=np.uint32) for k in PAF_COLS}
    h["is_rev"] = np.zeros(row_cap, dtype=np.uint8); h["mapq"] = np.zeros(row_cap, dtype=np.uint8)
    h["cg_off"] = np.zeros(row_cap + 1, dtype=np.uint32); h["cg_ops"] = np.zeros(op_cap, dtype=np.uint32)
    nops = C.c_uint64(0)
    fn.restype = C.c_int64
    fn.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64] + [u32p] * 4 + [u8p] + [u32p] * 6 + [u8p, u32p, u32p, C.POINTER(C.c_uint64)]
    n = fn(buf.ctypes.data if len(buf) else None, len(buf), row_cap, op_cap, _p(h["q_id"], u32p), _p(h["q_len"], u32p), _p(h["q_start"], u32p),
           _p(h["q_end"], u32p), _p(h["is_rev"], u8p), _p(h["t_id"], u32p), _p(h["t_len"], u32p), _p(h["t_start"], u32p), _p(h["t_end"], u32p),
           _p(h["n_match"], u32p), _p(h["n_block"], u32p), _p(h["mapq"], u8p), _p(h["cg_off"], u32p), _p(h["cg_ops"], u32p), C.byref(nops))
    if n < 0:
        return None, n
    for k in PAF_COLS + ("is_rev", "mapq"):
        h[k] = h[k][:n].copy()
    h["cg_off"] = h["cg_off"][: n + 1].copy()
    h["cg_ops"] = h["cg_ops"][: max(nops.value, 1)].copy()
    return h, n


def parse_paf(text):
    """K0 oracle (Longread.cpp:250-289, text side)."""
    return parse_paf_with(lib().oracle_parse_paf, text)
