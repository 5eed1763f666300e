"""CPU check of the product's per-read core (haslr_b200/csrc/k1_core.cuh, __host__ __device__) against the oracle
and the reference text, plus the restated libstdc++ std::sort against the real one on tie-heavy keys."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_io
import io_helpers
import oracle_ffi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "k1_host_check.cpp")
LIB = os.path.join(HERE, "native", "libk1test.so")
u8p, u32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_double))


@pytest.fixture(scope="module")
def k1():
    hdr = os.path.join(HERE, "..", "haslr_b200", "csrc", "k1_core.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    L.k1host_compact_lr.restype = C.c_int64
    L.k1host_compact_lr.argtypes = [u32p] * 8 + [u8p, u8p, u32p, u32p, u32p, C.c_uint32, f64p, C.c_double, C.c_double, C.c_double,
                                    C.c_uint32, C.c_uint32, C.c_void_p, u32p]
    L.k1host_sort_check.restype = C.c_int
    L.k1host_sort_check.argtypes = [u32p, u32p, C.c_uint32, u32p]
    return L


def run_k1(L, hits, read_off, mean_kmer, uniq_freq, min_aln_block=500):
    n_reads = len(read_off) - 1
    elems = np.zeros(max(len(hits["q_start"]), 1), dtype=oracle_ffi.CL_ELEM)
    off = np.zeros(n_reads + 1, dtype=np.uint32)
    p = lambda a, t: a.ctypes.data_as(t)
    n = L.k1host_compact_lr(*[p(hits[k], u32p) for k in ("q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block")],
                            p(hits["is_rev"], u8p), p(hits["mapq"], u8p), p(hits["cg_off"], u32p), p(hits["cg_ops"], u32p),
                            p(read_off, u32p), n_reads, p(mean_kmer, f64p), 0.85, uniq_freq, 0.15, min_aln_block, 55,
                            elems.ctypes.data, p(off, u32p))
    return elems[:n], off


def test_core_matches_oracle_and_reference(k1, oracle):
    g = golden_io.inputs()
    got, goff = run_k1(k1, g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"])
    ref, roff = oracle.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"])
    assert np.array_equal(goff, roff)
    assert got.tobytes() == ref.tobytes()
    assert io_helpers.format_compact(got, goff, g["hits"]) == golden_io.text("syn200k_compact_uniq.txt")


def test_core_other_thresholds(k1, oracle):
    g = golden_io.inputs()
    for blk in (100, 300, 900):
        got, goff = run_k1(k1, g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"], min_aln_block=blk)
        ref, roff = oracle.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"], min_aln_block=blk)
        assert np.array_equal(goff, roff) and got.tobytes() == ref.tobytes()


@pytest.mark.parametrize("n", [0, 1, 2, 15, 16, 17, 33, 100, 1000, 5000])
def test_restated_std_sort_matches_libstdcxx(k1, n):
    rng = np.random.default_rng(n)
    for keyspace in (2, 5, 50, 10**6):          # few distinct keys => many ties
        qe = rng.integers(0, keyspace, n).astype(np.uint32)
        qs = rng.integers(0, max(1, keyspace // 2), n).astype(np.uint32)
        if n == 0:
            qe = np.zeros(1, np.uint32); qs = np.zeros(1, np.uint32)
        bad = k1.k1host_sort_check(qe.ctypes.data_as(u32p), qs.ctypes.data_as(u32p), n, None)
        assert bad == 0
    # adversarial shapes: sorted, reversed, organ pipe (exercise the heapsort fallback bound)
    for arr in (np.arange(n), np.arange(n)[::-1], np.concatenate((np.arange(n // 2), np.arange(n - n // 2)[::-1]))):
        a = np.ascontiguousarray(arr, dtype=np.uint32)
        if n == 0:
            a = np.zeros(1, np.uint32)
        assert k1.k1host_sort_check(a.ctypes.data_as(u32p), a.ctypes.data_as(u32p), n, None) == 0


@pytest.mark.parametrize("seed", [1, 2])
def test_core_on_adversarial_hits(k1, oracle, seed):
    """tests/golden/k1adv_*: heavy overlaps, sort ties, repeated contigs — the product core against the oracle and the reference text."""
    a = golden_io.k1_adversarial(seed)
    hits, _ = oracle.parse_paf(a["paf"])
    read_off = np.searchsorted(hits["q_id"], np.arange(a["n_reads"] + 1), side="left").astype(np.uint32)
    uf = io_helpers.calc_uniq_freq(a["contig_len"], a["mean_kmer"])
    got, goff = run_k1(k1, hits, read_off, a["mean_kmer"], uf)
    ref, roff = oracle.compact_lr(hits, read_off, a["mean_kmer"], uf)
    assert np.array_equal(goff, roff) and got.tobytes() == ref.tobytes()
    assert io_helpers.format_compact(got, goff, hits) == a["compact"]
