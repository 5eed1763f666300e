"""ctypes binding of the C ABI declared in include/haslr_b200.h (one entry per exported function)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
u8p, u32p, u64p, i32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32, C.c_double))

HGPU_OK, HGPU_E_INVALID, HGPU_E_CUDA, HGPU_E_NOMEM, HGPU_E_NOSPACE, HGPU_E_UNSUPPORTED, HGPU_E_INTERNAL = 0, -1, -2, -3, -4, -5, -6

# every symbol include/haslr_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "hgpu_create", "hgpu_destroy", "hgpu_set_stream", "hgpu_strerror", "hgpu_last_error", "hgpu_abi_version",
    "hgpu_launch_count", "hgpu_compact_lr", "hgpu_backbone_edges", "hgpu_poa_batch", "hgpu_poa_batch_dev",
    "hgpu_poa_fetch", "hgpu_poa_get_stats", "hgpu_poa_set_timing", "hgpu_poa_configure", "hgpu_poa_debug", "hgpu_edge_coords",
    "hgpu_paf_tokenize", "hgpu_paf_fetch", "hgpu_hits_group", "hgpu_compact_lr_dev", "hgpu_backbone_edges_dev", "hgpu_edge_coords_dev",
    "hgpu_get_stage_stats", "hgpu_set_timing", "hgpu_host_staging",
)


class HgpuError(RuntimeError):
    def __init__(self, code, detail):
        super().__init__(f"hgpu error {code}: {detail}")
        self.code = code


class HitsT(C.Structure):
    _fields_ = [("n_hits", C.c_uint32)] + [(n, u32p) for n in
                ("q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block")] + \
               [("is_rev", u8p), ("mapq", u8p), ("cg_off", u32p), ("cg_ops", u32p)]


class K1Params(C.Structure):
    _fields_ = [("min_aln_sim", C.c_double), ("uniq_freq", C.c_double), ("max_uniq_dev", C.c_double),
                ("min_aln_block", C.c_uint32), ("min_aln_mapq", C.c_uint32)]


class PoaStats(C.Structure):
    _fields_ = [("cells", C.c_uint64), ("cells_padded", C.c_uint64), ("alignments", C.c_uint64),
                ("alignments_i32", C.c_uint64), ("bases_in", C.c_uint64), ("bases_out", C.c_uint64),
                ("dp_launches", C.c_uint64), ("update_launches", C.c_uint64), ("other_launches", C.c_uint64),
                ("ms_dp", C.c_float), ("ms_update", C.c_float), ("ms_other", C.c_float), ("arena_bytes", C.c_uint64),
                ("alignments_rel16", C.c_uint64)]


class StageStats(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_k0", "ms_k1", "ms_k2", "ms_k4")] + \
               [(n, C.c_uint32) for n in ("launches_k0", "launches_k1", "launches_k2", "launches_k4")] + \
               [(n, C.c_uint64) for n in ("k0_text_bytes", "k0_rows", "k0_ops", "k1_hits", "k1_reads", "k1_elems", "k2_pairs", "k2_entries",
                                          "k4_edges", "k4_supports", "k4_runs", "h2d_bytes", "d2h_bytes")]


class DbgSizes(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_edges", C.c_uint32), ("aln_len", C.c_uint32), ("L", C.c_uint32)]


CL_ELEM = np.dtype([(n, "<u4") for n in ("hit", "q_start", "q_end", "t_start", "t_end", "n_match", "n_block",
                                         "cg_lo", "cg_lo_len", "cg_hi", "cg_hi_len")])
EDGE_SUPP = np.dtype([("lr_id_strand", "<u4"), ("cmp_head", "<u4"), ("cmp_tail", "<u4")])
EDGE_COORD = np.dtype([(n, "<u4") for n in ("int1_lo", "int1_hi", "int2_lo", "int2_hi", "c1", "c2", "n_best", "n_cns")])
SUPP_COORD = np.dtype([("lr_start", "<i8"), ("lr_end", "<i8"), ("lr_strand", "<u4"), ("in_best", "<u4")])


def lib_path():
    # HASLR_B200_LIB: developer override to A/B-test another build of the same library
    return os.environ.get("HASLR_B200_LIB") or os.path.join(_HERE, "libhaslr_b200.so")


_lib = None


def load():
    """Load the CUDA library. Raises if it has not been built (python __graft_entry__.py / make)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: build it with `make` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(p)
    L.hgpu_create.restype = C.c_int; L.hgpu_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.hgpu_destroy.restype = None; L.hgpu_destroy.argtypes = [C.c_void_p]
    L.hgpu_set_stream.restype = C.c_int; L.hgpu_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.hgpu_strerror.restype = C.c_char_p; L.hgpu_strerror.argtypes = [C.c_int]
    L.hgpu_last_error.restype = C.c_char_p; L.hgpu_last_error.argtypes = [C.c_void_p]
    L.hgpu_abi_version.restype = C.c_int; L.hgpu_abi_version.argtypes = []
    L.hgpu_launch_count.restype = C.c_uint64; L.hgpu_launch_count.argtypes = [C.c_void_p]
    L.hgpu_host_staging.restype = C.c_int; L.hgpu_host_staging.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]
    L.hgpu_compact_lr.restype = C.c_int
    L.hgpu_compact_lr.argtypes = [C.c_void_p, C.POINTER(HitsT), u32p, C.c_uint32, f64p, C.c_uint32, C.POINTER(K1Params),
                                  C.c_void_p, u32p, u64p]
    L.hgpu_backbone_edges.restype = C.c_int
    L.hgpu_backbone_edges.argtypes = [C.c_void_p, u32p, u8p, u32p, C.c_uint32, C.c_uint32, u64p, u32p, C.c_void_p, u8p, u64p]
    poa_sig = [C.c_void_p, C.c_void_p, u64p, u32p, C.c_uint32, C.c_int8, C.c_int8, C.c_int8, C.c_uint32,
               C.c_void_p, C.c_uint64, u64p, u32p]
    L.hgpu_poa_batch.restype = C.c_int; L.hgpu_poa_batch.argtypes = poa_sig
    L.hgpu_poa_batch_dev.restype = C.c_int; L.hgpu_poa_batch_dev.argtypes = poa_sig
    L.hgpu_poa_fetch.restype = C.c_int; L.hgpu_poa_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.hgpu_poa_get_stats.restype = C.c_int; L.hgpu_poa_get_stats.argtypes = [C.c_void_p, C.POINTER(PoaStats)]
    L.hgpu_poa_set_timing.restype = C.c_int; L.hgpu_poa_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.hgpu_poa_configure.restype = C.c_int; L.hgpu_poa_configure.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    L.hgpu_poa_debug.restype = C.c_int
    L.hgpu_poa_debug.argtypes = [C.c_void_p, u8p, u64p, C.c_uint32, C.c_uint32, C.c_int8, C.c_int8, C.c_int8, C.c_int, C.c_int,
                                 i32p, C.c_uint64, i32p, i32p, C.c_uint32, u32p, u8p, u32p, u32p, u32p,
                                 C.c_uint32, C.c_uint32, C.POINTER(DbgSizes)]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


class Context:
    """One hgpu context = one GPU. Mirrors the reference's stage functions (see include/haslr_b200.h)."""

    def __init__(self, device=-1):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.hgpu_create(device, C.byref(h))
        if rc != 0:
            raise HgpuError(rc, self.L.hgpu_strerror(rc).decode() + " (hgpu_create: a B200-class GPU is required; no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.hgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise HgpuError(rc, self.L.hgpu_last_error(self.h).decode())

    def set_stream(self, stream_ptr):
        self._check(self.L.hgpu_set_stream(self.h, C.c_void_p(stream_ptr)))

    def launch_count(self):
        return int(self.L.hgpu_launch_count(self.h))

    def host_staging(self, which, nbytes):
        """Address of the context's page-locked staging buffer `which` (0 / 1), at least nbytes long (hgpu_host_staging)."""
        p = C.c_void_p()
        self._check(self.L.hgpu_host_staging(self.h, which, nbytes, C.byref(p)))
        return p.value

    # ---- (iii) batched POA -------------------------------------------------------------------------------
    def poa_batch(self, bases, seg_off, edge_seg_off, match=5, mismatch=-4, gap=-8, band=0, out=None):
        """Host buffers in, host buffers out (copies inside). Returns (cons uint8[], cons_off uint64[n+1], status uint32[n])."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        seg_off = np.ascontiguousarray(seg_off, dtype=np.uint64)
        edge_seg_off = np.ascontiguousarray(edge_seg_off, dtype=np.uint32)
        n = len(edge_seg_off) - 1
        cap = int(seg_off[-1]) + 64 if out is None else len(out)
        if out is None:
            out = np.empty(cap, dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.uint32)
        rc = self.L.hgpu_poa_batch(self.h, bases.ctypes.data, _p(seg_off, u64p), _p(edge_seg_off, u32p), n, match, mismatch, gap, band,
                                   out.ctypes.data, cap, _p(off, u64p), _p(status, u32p))
        self._check(rc)
        return out[: int(off[n])], off, status[:n]

    def poa_batch_dev(self, d_bases_ptr, seg_off, edge_seg_off, d_out_ptr, out_cap, match=5, mismatch=-4, gap=-8, band=0):
        """Bases and consensus already/still on the device (raw pointers). Returns (cons_off, status)."""
        seg_off = np.ascontiguousarray(seg_off, dtype=np.uint64)
        edge_seg_off = np.ascontiguousarray(edge_seg_off, dtype=np.uint32)
        n = len(edge_seg_off) - 1
        off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.uint32)
        rc = self.L.hgpu_poa_batch_dev(self.h, C.c_void_p(d_bases_ptr), _p(seg_off, u64p), _p(edge_seg_off, u32p), n, match, mismatch, gap,
                                       band, C.c_void_p(d_out_ptr), out_cap, _p(off, u64p), _p(status, u32p))
        self._check(rc)
        return off, status[:n]

    def poa_stats(self):
        s = PoaStats()
        self._check(self.L.hgpu_poa_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in PoaStats._fields_}

    def poa_set_timing(self, on):
        self._check(self.L.hgpu_poa_set_timing(self.h, int(on)))

    def poa_configure(self, arena_bytes=0, max_warps=0):
        self._check(self.L.hgpu_poa_configure(self.h, arena_bytes, max_warps))

    def poa_debug(self, bases, seg_off, n_prior, match=5, mismatch=-4, gap=-8, force_i32=0, want_H=True):
        """force_i32: 0 = the cell encoding the batch path picks, 1 = int32, 2 = row-relative int16 (REL16).
        Graph after n_prior segments (rank order) + score matrix / alignment of the next one; same dict as the oracle's."""
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
        poff = np.zeros(cap_n + 1, dtype=np.uint32); pn = np.zeros(2 * cap_n + n_segs + 8, dtype=np.uint32); pw = np.zeros_like(pn)
        sz = DbgSizes()
        rc = self.L.hgpu_poa_debug(self.h, _p(bases, u8p), _p(seg_off, u64p), n_segs, n_prior, match, mismatch, gap, int(force_i32), 0,
                                   _p(H, i32p) if want_H else None, len(H) if want_H else 0, _p(an, i32p), _p(ap, i32p), aln_cap,
                                   _p(r2n, u32p), _p(code, u8p), _p(poff, u32p), _p(pn, u32p), _p(pw, u32p), cap_n, len(pn), C.byref(sz))
        self._check(rc)
        V, L = sz.n_nodes, sz.L
        return dict(V=V, L=L, H=H[: (V + 1) * (L + 1)].reshape(V + 1, L + 1) if want_H else None,
                    aln_node=an[: sz.aln_len].copy(), aln_pos=ap[: sz.aln_len].copy(), rank2node=r2n[:V].copy(),
                    code=code[:V].copy(), pred_off=poff[: V + 1].copy(), pred_node=pn[: int(poff[V])].copy(),
                    pred_weight=pw[: int(poff[V])].copy())


# ---- (i) / (ii): methods attached to Context ----------------------------------------------------------------
def _hits_struct(h):
    s = HitsT()
    s.n_hits = len(h["q_start"])
    keep = []
    for n in ("q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block", "cg_off", "cg_ops"):
        a = np.ascontiguousarray(h[n], dtype=np.uint32); keep.append(a)
        setattr(s, n, _p(a, u32p))
    for n in ("is_rev", "mapq"):
        a = np.ascontiguousarray(h[n], dtype=np.uint8); keep.append(a)
        setattr(s, n, _p(a, u8p))
    return s, keep


def _compact_lr(self, hits, read_off, mean_kmer, uniq_freq, min_aln_block=500, min_aln_sim=0.85, min_aln_mapq=55, max_uniq_dev=0.15):
    """build_compact_longreads (Longread.hpp:89) incl. the load filters / sort / overlap fix. Returns (elems CL_ELEM[], read_off)."""
    hs, keep = _hits_struct(hits)
    read_off = np.ascontiguousarray(read_off, dtype=np.uint32)
    mean_kmer = np.ascontiguousarray(mean_kmer, dtype=np.float64)
    n_reads = len(read_off) - 1
    prm = K1Params(min_aln_sim, uniq_freq, max_uniq_dev, min_aln_block, min_aln_mapq)
    elems = np.zeros(max(hs.n_hits, 1), dtype=CL_ELEM)
    out_off = np.zeros(n_reads + 1, dtype=np.uint32)
    n = C.c_uint64(0)
    rc = self.L.hgpu_compact_lr(self.h, C.byref(hs), _p(read_off, u32p), n_reads, _p(mean_kmer, f64p), len(mean_kmer), C.byref(prm),
                                elems.ctypes.data, _p(out_off, u32p), C.byref(n))
    self._check(rc)
    return elems[: n.value].copy(), out_off


def _backbone_edges(self, cl_tid, cl_rev, cl_read_off, min_edge_sup=3):
    """bbg_build_graph + the weak-edge rule (Backbone_graph.hpp:60,63). Returns (key64[], supp_off[], supp EDGE_SUPP[], keep[])."""
    cl_tid = np.ascontiguousarray(cl_tid, dtype=np.uint32)
    cl_rev = np.ascontiguousarray(cl_rev, dtype=np.uint8)
    cl_read_off = np.ascontiguousarray(cl_read_off, dtype=np.uint32)
    n_reads = len(cl_read_off) - 1
    cnt = np.diff(cl_read_off.astype(np.int64))
    n_pairs = int(np.maximum(cnt - 1, 0).sum())
    cap = max(2 * n_pairs, 1)
    key = np.zeros(cap, dtype=np.uint64); soff = np.zeros(cap + 1, dtype=np.uint32)
    supp = np.zeros(cap, dtype=EDGE_SUPP); keep = np.zeros(cap, dtype=np.uint8)
    n = C.c_uint64(0)
    rc = self.L.hgpu_backbone_edges(self.h, _p(cl_tid, u32p), _p(cl_rev, u8p), _p(cl_read_off, u32p), n_reads, min_edge_sup,
                                    _p(key, u64p), _p(soff, u32p), supp.ctypes.data, _p(keep, u8p), C.byref(n))
    self._check(rc)
    n = n.value
    return key[:n].copy(), soff[: n + 1].copy(), supp[: int(soff[n])].copy(), keep[:n].copy()


Context.compact_lr = _compact_lr
Context.backbone_edges = _backbone_edges


def _edge_coords(self, edge_rev, supp_off, supp, elems, cl_read_off, read_len, hits):
    """hgpu_edge_coords: per edge (rev1 | rev2 << 1) and its supports -> (EDGE_COORD[n_edges], SUPP_COORD[n_supp])."""
    edge_rev = np.ascontiguousarray(edge_rev, dtype=np.uint8)
    supp_off = np.ascontiguousarray(supp_off, dtype=np.uint32)
    supp = np.ascontiguousarray(supp, dtype=EDGE_SUPP)
    elems = np.ascontiguousarray(elems, dtype=CL_ELEM)
    cl_read_off = np.ascontiguousarray(cl_read_off, dtype=np.uint32)
    read_len = np.ascontiguousarray(read_len, dtype=np.uint32)
    is_rev = np.ascontiguousarray(hits["is_rev"], dtype=np.uint8)
    cg_off = np.ascontiguousarray(hits["cg_off"], dtype=np.uint32)
    cg_ops = np.ascontiguousarray(hits["cg_ops"], dtype=np.uint32)
    n = len(edge_rev)
    oe = np.zeros(max(n, 1), dtype=EDGE_COORD); os_ = np.zeros(max(len(supp), 1), dtype=SUPP_COORD)
    self.L.hgpu_edge_coords.restype = C.c_int
    self.L.hgpu_edge_coords.argtypes = [C.c_void_p, C.c_uint32, u8p, u32p, C.c_void_p, C.c_void_p, u32p, C.c_uint32, u32p, u8p, u32p, u32p,
                                        C.c_uint32, C.c_void_p, C.c_void_p]
    self._check(self.L.hgpu_edge_coords(self.h, n, _p(edge_rev, u8p), _p(supp_off, u32p), supp.ctypes.data, elems.ctypes.data,
                                        _p(cl_read_off, u32p), len(cl_read_off) - 1, _p(read_len, u32p), _p(is_rev, u8p), _p(cg_off, u32p),
                                        _p(cg_ops, u32p), len(is_rev), oe.ctypes.data, os_.ctypes.data))
    return oe[:n], os_[: len(supp)]


Context.edge_coords = _edge_coords


def _parse_paf(self, text):
    """hgpu_paf_tokenize + hgpu_paf_fetch: PAF bytes -> hits dict (same keys as tests/io_helpers.parse_paf)."""
    buf = np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text, dtype=np.uint8)
    nr, no = C.c_uint64(0), C.c_uint64(0)
    self.L.hgpu_paf_tokenize.restype = C.c_int
    self.L.hgpu_paf_tokenize.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    self._check(self.L.hgpu_paf_tokenize(self.h, buf.ctypes.data if len(buf) else None, len(buf), C.byref(nr), C.byref(no)))
    n, m = nr.value, no.value
    names = ("q_id", "q_len", "q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block")
    h = {k: np.zeros(max(n, 1), dtype=np.uint32) for k in names}
    h["is_rev"] = np.zeros(max(n, 1), dtype=np.uint8); h["mapq"] = np.zeros(max(n, 1), dtype=np.uint8)
    h["cg_off"] = np.zeros(n + 1, dtype=np.uint32); h["cg_ops"] = np.zeros(max(m, 1), dtype=np.uint32)
    self.L.hgpu_paf_fetch.restype = C.c_int
    self.L.hgpu_paf_fetch.argtypes = [C.c_void_p] + [u32p] * 4 + [u8p] + [u32p] * 6 + [u8p, u32p, u32p]
    self._check(self.L.hgpu_paf_fetch(self.h, _p(h["q_id"], u32p), _p(h["q_len"], u32p), _p(h["q_start"], u32p), _p(h["q_end"], u32p),
                                      _p(h["is_rev"], u8p), _p(h["t_id"], u32p), _p(h["t_len"], u32p), _p(h["t_start"], u32p), _p(h["t_end"], u32p),
                                      _p(h["n_match"], u32p), _p(h["n_block"], u32p), _p(h["mapq"], u8p), _p(h["cg_off"], u32p), _p(h["cg_ops"], u32p)))
    for k in names + ("is_rev", "mapq"):
        h[k] = h[k][:n]
    if m == 0:
        h["cg_ops"] = np.zeros(1, dtype=np.uint32)       # io_helpers.parse_paf keeps one dummy word so pointers stay valid
    return h


Context.parse_paf = _parse_paf


# ---- device-resident stages: the hit table hgpu_paf_tokenize made stays on the device -------------------------------
def _tokenize(self, text):
    """hgpu_paf_tokenize only: the hit table stays on the device. Returns (n_rows, n_ops)."""
    buf = np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text, dtype=np.uint8)
    nr, no = C.c_uint64(0), C.c_uint64(0)
    self.L.hgpu_paf_tokenize.restype = C.c_int
    self.L.hgpu_paf_tokenize.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    self._check(self.L.hgpu_paf_tokenize(self.h, buf.ctypes.data if len(buf) else None, len(buf), C.byref(nr), C.byref(no)))
    return nr.value, no.value


def _hits_group(self, n_reads):
    off = np.zeros(n_reads + 1, dtype=np.uint32)
    self.L.hgpu_hits_group.restype = C.c_int
    self.L.hgpu_hits_group.argtypes = [C.c_void_p, C.c_uint32, u32p]
    self._check(self.L.hgpu_hits_group(self.h, n_reads, _p(off, u32p)))
    return off


def _compact_lr_dev(self, n_reads, n_hits, mean_kmer, uniq_freq, min_aln_block=500, min_aln_sim=0.85, min_aln_mapq=55, max_uniq_dev=0.15):
    """Returns (elems CL_ELEM[], tid uint32[], rev uint8[], read_off uint32[n_reads+1]); the elements also stay on the device."""
    mean_kmer = np.ascontiguousarray(mean_kmer, dtype=np.float64)
    prm = K1Params(min_aln_sim, uniq_freq, max_uniq_dev, min_aln_block, min_aln_mapq)
    elems = np.zeros(max(n_hits, 1), dtype=CL_ELEM); tid = np.zeros(max(n_hits, 1), dtype=np.uint32); rev = np.zeros(max(n_hits, 1), dtype=np.uint8)
    out_off = np.zeros(n_reads + 1, dtype=np.uint32)
    n = C.c_uint64(0)
    self.L.hgpu_compact_lr_dev.restype = C.c_int
    self.L.hgpu_compact_lr_dev.argtypes = [C.c_void_p, C.c_uint32, f64p, C.c_uint32, C.POINTER(K1Params), C.c_void_p, u32p, u8p, u32p, u64p]
    self._check(self.L.hgpu_compact_lr_dev(self.h, n_reads, _p(mean_kmer, f64p), len(mean_kmer), C.byref(prm), elems.ctypes.data, _p(tid, u32p),
                                           _p(rev, u8p), _p(out_off, u32p), C.byref(n)))
    n = n.value
    return elems[:n].copy(), tid[:n].copy(), rev[:n].copy(), out_off


def _backbone_edges_dev(self, cl_read_off, min_edge_sup=3):
    cnt = np.diff(np.asarray(cl_read_off).astype(np.int64))
    cap = max(2 * int(np.maximum(cnt - 1, 0).sum()), 1)
    key = np.zeros(cap, dtype=np.uint64); soff = np.zeros(cap + 1, dtype=np.uint32)
    supp = np.zeros(cap, dtype=EDGE_SUPP); keep = np.zeros(cap, dtype=np.uint8)
    n = C.c_uint64(0)
    self.L.hgpu_backbone_edges_dev.restype = C.c_int
    self.L.hgpu_backbone_edges_dev.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, u64p, u32p, C.c_void_p, u8p, u64p]
    self._check(self.L.hgpu_backbone_edges_dev(self.h, min_edge_sup, cap, _p(key, u64p), _p(soff, u32p), supp.ctypes.data, _p(keep, u8p), C.byref(n)))
    n = n.value
    return key[:n].copy(), soff[: n + 1].copy(), supp[: int(soff[n])].copy(), keep[:n].copy()


def _edge_coords_dev(self, edge_rev, supp_off, supp, read_len):
    edge_rev = np.ascontiguousarray(edge_rev, dtype=np.uint8)
    supp_off = np.ascontiguousarray(supp_off, dtype=np.uint32)
    supp = np.ascontiguousarray(supp, dtype=EDGE_SUPP)
    read_len = np.ascontiguousarray(read_len, dtype=np.uint32)
    n = len(edge_rev)
    oe = np.zeros(max(n, 1), dtype=EDGE_COORD); os_ = np.zeros(max(len(supp), 1), dtype=SUPP_COORD)
    self.L.hgpu_edge_coords_dev.restype = C.c_int
    self.L.hgpu_edge_coords_dev.argtypes = [C.c_void_p, C.c_uint32, u8p, u32p, C.c_void_p, u32p, C.c_uint32, C.c_void_p, C.c_void_p]
    self._check(self.L.hgpu_edge_coords_dev(self.h, n, _p(edge_rev, u8p), _p(supp_off, u32p), supp.ctypes.data, _p(read_len, u32p), len(read_len),
                                            oe.ctypes.data, os_.ctypes.data))
    return oe[:n], os_[: len(supp)]


def _stage_stats(self):
    s = StageStats()
    self.L.hgpu_get_stage_stats.restype = C.c_int
    self.L.hgpu_get_stage_stats.argtypes = [C.c_void_p, C.POINTER(StageStats)]
    self._check(self.L.hgpu_get_stage_stats(self.h, C.byref(s)))
    return {k: getattr(s, k) for k, _ in StageStats._fields_}


def _set_timing(self, on):
    self.L.hgpu_set_timing.restype = C.c_int
    self.L.hgpu_set_timing.argtypes = [C.c_void_p, C.c_int]
    self._check(self.L.hgpu_set_timing(self.h, int(on)))


Context.tokenize = _tokenize
Context.hits_group = _hits_group
Context.compact_lr_dev = _compact_lr_dev
Context.backbone_edges_dev = _backbone_edges_dev
Context.edge_coords_dev = _edge_coords_dev
Context.stage_stats = _stage_stats
Context.set_timing = _set_timing
