"""CPU check of the drop-in binary's stitching code (haslr_b200/host/assemble.cpp: edge enumeration in asm_get_next_edge
order, simple-path extraction, assemble_path / write_assembly — Assemble.cpp:365-434,607-810,1045-1077) against the
assembly the reference binary wrote for the golden dataset. Edge table and edge coordinates come from the oracle, the
consensus strings from the golden POA fixture (what the reference logged)."""
import ctypes as C
import gzip
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_io
import io_helpers
import oracle_ffi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PROD = os.path.join(ROOT, "haslr_b200")
SRC = os.path.join(HERE, "native", "asm_host_check.cpp")
LIB = os.path.join(HERE, "native", "libasmtest.so")
u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def asm():
    if not os.path.exists(os.path.join(PROD, "libhaslr_b200.so")):
        pytest.skip("product library not built")
    srcs = [SRC] + [os.path.join(PROD, "host", f) for f in ("assemble.cpp", "bbg.cpp", "io.cpp")]
    deps = srcs + [os.path.join(PROD, "host", "haslr.hpp")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", LIB] + srcs + ["-L" + PROD, "-lhaslr_b200", "-Wl,-rpath," + PROD, "-lz", "-lpthread"], check=True)
    return C.CDLL(LIB)


def test_assembly_matches_reference(asm, oracle, tmp_path):
    g = golden_io.inputs()
    with tempfile.TemporaryDirectory() as tmp:          # the golden dataset's contigs (tests/golden/make_golden.py: 200 kb, 600 reads, seed 7)
        subprocess.run([oracle_ffi.GEN_BIN, tmp, "200000", "600", "8000", "7"], check=True, stdout=subprocess.DEVNULL)
        contigs = io_helpers.load_fasta(os.path.join(tmp, "contigs.fa"))
    assert [len(c) for c in contigs] == list(g["contig_len"])
    seq = b"".join(contigs)
    coff = np.concatenate(([0], np.cumsum([len(c) for c in contigs]))).astype(np.uint64)
    elems, off = oracle.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"])
    key, soff, supp, _ = oracle.backbone_edges(g["hits"]["t_id"][elems["hit"]], g["hits"]["is_rev"][elems["hit"]], off, 3)
    p = lambda a, t: a.ctypes.data_as(t)
    asm.asmhost_prepare.argtypes = [C.c_uint32, C.c_char_p, u64p, C.c_uint64, u64p, u32p, C.c_void_p, C.c_uint32, C.c_char_p]
    n = asm.asmhost_prepare(len(contigs), seq, p(coff, u64p), len(key), p(key, u64p), p(soff, u32p), supp.ctypes.data, 3, str(tmp_path).encode())
    gold = golden_io.coords()
    assert n == len(gold) == 120
    e4 = np.zeros(4 * n, dtype=np.uint32); eso = np.zeros(n + 1, dtype=np.uint32); esupp = np.zeros(len(supp), dtype=oracle_ffi.EDGE_SUPP)
    asm.asmhost_edges.argtypes = [u32p, u32p, C.c_void_p, C.c_uint32]
    ns = asm.asmhost_edges(p(e4, u32p), p(eso, u32p), esupp.ctypes.data, len(esupp))
    assert ns >= 0
    e4 = e4.reshape(n, 4)
    assert [tuple(r) for r in e4.tolist()] == [x["edge"] for x in gold]                  # asm_get_next_edge order
    oe, _ = oracle.edge_coords((e4[:, 1] | (e4[:, 3] << 1)).astype(np.uint8), eso, esupp[:ns], elems, off, golden_io.read_len(), g["hits"])
    lens = g["contig_len"].astype(np.int64)
    he = np.where(oe["n_cns"] > 0, oe["c1"], np.where(e4[:, 1] == 0, lens[e4[:, 0]] - 1, 0)).astype(np.uint32)     # Assemble.cpp:339-361
    tb = np.where(oe["n_cns"] > 0, oe["c2"], np.where(e4[:, 3] == 0, 0, lens[e4[:, 2]] - 1)).astype(np.uint32)
    cons_by_label = {lab: c for lab, _, c in golden_io.poa_edges()}
    cons = [cons_by_label["%d:%s -> %d:%s" % (a, "+-"[b], c, "+-"[d])] for a, b, c, d in e4.tolist()]
    cns_off = np.concatenate(([0], np.cumsum([len(c) for c in cons]))).astype(np.uint64)
    out = tmp_path / "out"; out.mkdir()
    asm.asmhost_finish.argtypes = [u32p, u32p, u32p, C.c_char_p, u64p, C.c_char_p]
    ncns = np.ascontiguousarray(oe["n_cns"], dtype=np.uint32)
    assert asm.asmhost_finish(p(he, u32p), p(tb, u32p), p(ncns, u32p), b"".join(cons), p(cns_off, u64p), str(out).encode()) == 0
    with open(out / "asm.final.fa", "rb") as f, gzip.open(os.path.join(golden_io.GOLD, "syn200k_asm.final.fa.gz"), "rb") as gz:
        assert f.read() == gz.read()
    with open(out / "asm.final.ann") as f:
        assert f.read() == golden_io.text("syn200k_asm.final.ann")


def test_segments_match_what_the_reference_fed_to_spoa(asm, oracle):
    """The product's segment extraction (read or reverse complement from spos, uint32 length arithmetic) on the cns_supp lists the
    oracle's edge coordinates give, against the segments the reference binary logged for the golden dataset (log_consensus.txt)."""
    g = golden_io.inputs()
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([oracle_ffi.GEN_BIN, tmp, "200000", "600", "8000", "7"], check=True, stdout=subprocess.DEVNULL)
        reads = [bytes(c if c in b"ACGT" else 65 for c in r.upper()) for r in io_helpers.load_fasta(os.path.join(tmp, "reads.fa"))]
    assert [len(r) for r in reads] == list(golden_io.read_len())
    ci = golden_io.coord_inputs(oracle)
    oe, os_ = oracle.edge_coords(ci["edge_rev"], ci["supp_off"], ci["supp"], ci["elems"], ci["cl_off"], ci["read_len"], g["hits"])
    want = golden_io.poa_edges()
    assert len(want) == len(ci["gold"]) == 120
    cns, per_edge = [], []
    for e in range(120):
        k = 0
        for i in range(int(ci["supp_off"][e]), int(ci["supp_off"][e + 1])):
            o = os_[i]
            if o["in_best"] and o["lr_start"] != -1 and o["lr_end"] != -1:
                cns.append((int(ci["supp"][i]["lr_id_strand"]) & 0x7FFFFFFF, int(o["lr_strand"]), int(o["lr_start"]) + 1, int(o["lr_end"]) - 1)); k += 1
        per_edge.append(k)
    cns4 = np.array(cns, dtype=np.uint32).reshape(-1)
    roff = np.concatenate(([0], np.cumsum([len(r) for r in reads]))).astype(np.uint64)
    cap = int(sum(len(s) for _, segs, _ in want for s in segs)) + 1024
    out = C.create_string_buffer(cap); ooff = np.zeros(len(cns) + 1, dtype=np.uint64)
    asm.asmhost_segments.restype = C.c_longlong
    asm.asmhost_segments.argtypes = [C.c_char_p, u64p, C.c_uint32, u32p, C.c_uint32, C.c_char_p, C.c_uint64, u64p]
    n = asm.asmhost_segments(b"".join(reads), roff.ctypes.data_as(u64p), len(reads), cns4.ctypes.data_as(u32p), len(cns), out, cap, ooff.ctypes.data_as(u64p))
    assert n >= 0
    k = 0
    for e, (label, segs, _) in enumerate(want):
        assert per_edge[e] == len(segs), label
        for s in segs:
            assert out.raw[int(ooff[k]): int(ooff[k + 1])] == s, (label, k)
            k += 1
    assert k == len(cns)
