"""CPU check of the product's POA graph code (haslr_b200/csrc/poa_graph.cuh is __host__ __device__) against the oracle.

tests/native/graph_host_check.cpp drives the product's add_alignment / topological sort / per-rank DP records /
consensus with a scalar graph-NW of its own; the consensus and the final graph must equal the oracle's bit for bit.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "graph_host_check.cpp")
LIB = os.path.join(HERE, "native", "libgraphtest.so")
u8p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64))


@pytest.fixture(scope="module")
def gt():
    hdr = os.path.join(HERE, "..", "haslr_b200", "csrc", "poa_graph.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    L.graphtest_poa.restype = C.c_int
    L.graphtest_poa.argtypes = [u8p, u64p, C.c_uint32, C.c_int, C.c_int, C.c_int, u8p, C.c_uint32, u32p, u32p, u32p, u32p, u32p, C.c_uint32]
    return L


def run_host(gt, bases, seg_off, scores=(5, -4, -8)):
    n_segs = len(seg_off) - 1
    cap = int(seg_off[-1] - seg_off[0]) + 64
    out = np.zeros(cap, dtype=np.uint8)
    nn = C.c_uint32(0)
    r2n = np.zeros(cap, dtype=np.uint32); poff = np.zeros(cap + 1, dtype=np.uint32)
    pn = np.zeros(2 * cap, dtype=np.uint32); pw = np.zeros(2 * cap, dtype=np.uint32)
    b = np.ascontiguousarray(bases)
    so = np.ascontiguousarray(seg_off, dtype=np.uint64)
    n = gt.graphtest_poa(b.ctypes.data_as(u8p), so.ctypes.data_as(u64p), n_segs, *scores, out.ctypes.data_as(u8p), cap,
                         C.byref(nn), r2n.ctypes.data_as(u32p), poff.ctypes.data_as(u32p), pn.ctypes.data_as(u32p),
                         pw.ctypes.data_as(u32p), cap)
    assert n >= 0, f"graphtest_poa failed with {n}"
    V = nn.value
    return out[:n].tobytes(), dict(V=V, rank2node=r2n[:V], pred_off=poff[: V + 1], pred_node=pn[: poff[V]], pred_weight=pw[: poff[V]])


@pytest.mark.parametrize("seed,depth,length,err", [
    (1, 6, 300, (0.04, 0.03, 0.02)),
    (2, 12, 200, (0.08, 0.06, 0.04)),
    (3, 3, 700, (0.04, 0.03, 0.02)),
    (4, 20, 120, (0.10, 0.08, 0.06)),
])
def test_graph_code_matches_oracle(gt, oracle, seed, depth, length, err):
    bases, seg_off, eso, _ = synth.poa_batch(seed, 6, depth=depth, length=length, err=err, length_jitter=0.2)
    cons, off, _, nodes = oracle.poa_batch(bases, seg_off, eso)
    for e in range(len(eso) - 1):
        so = seg_off[eso[e]: eso[e + 1] + 1]
        got, g = run_host(gt, bases, so)
        assert got == cons[int(off[e]): int(off[e + 1])].tobytes()
        assert g["V"] == nodes[e]
        # final graph, rank order: same topological order, same in-edge order and weights
        ref = oracle.poa_debug(bases[int(so[0]): int(so[-1])], so - so[0], n_prior=len(so) - 1, want_H=False)
        assert np.array_equal(g["rank2node"], ref["rank2node"])
        assert np.array_equal(g["pred_off"], ref["pred_off"])
        assert np.array_equal(g["pred_node"], ref["pred_node"])
        assert np.array_equal(g["pred_weight"], ref["pred_weight"])


def test_graph_code_edge_cases(gt, oracle):
    cases = [
        [b"ACGT"],                                   # single segment: consensus is the segment
        [b"A", b"A", b"A"],
        [b"ACGTACGT", b"", b"ACGTACGT"],              # empty segments are skipped (Assemble.cpp:537)
        [b"AAAAAAAAAA", b"TTTTTTTTTT", b"AAAAAAAAAA"],
        [b"ACGTNNACGT", b"ACGTAAACGT", b"acgtaaacgt"],  # non-ACGT and lower case fold to A / upper (Compressed_sequence.cpp:57)
        [b"ACGT" * 30, b"ACGT" * 10, b"ACGT" * 50],   # very different lengths
        [b"GATTACA", b"GATACA", b"GATTTACA", b"CATTACA", b"GATTACAT", b"TGATTACA"],
    ]
    for segs in cases:
        bases, seg_off, eso = synth.from_strings([segs])
        cons, off, _, _ = oracle.poa_batch(bases, seg_off, eso)
        got, _ = run_host(gt, bases, seg_off)
        assert got == cons.tobytes(), segs


@pytest.mark.parametrize("seed,depth,length,err", [
    (5, 28, 800, (0.04, 0.03, 0.02)),      # the config 2 edge shape, shortened
    (8, 30, 1500, (0.10, 0.08, 0.06)),     # 24 % read error: wide graphs, long branches
    (9, 6, 1500, (0.04, 0.03, 0.02)),      # config 3
])
def test_parallel_toposort_formulation_is_exact(gt, seed, depth, length, err):
    """graphtest_poa re-derives every topological order of the run twice more - by the claim rule (smallest root id that reaches
    a node, walks on disjoint node sets: -5 on a difference) and in the memory layout of the CUDA function w_toposort_claims
    (-6) - and fails the call if either differs from SPOA's serial walk. The device layout must never have to decline, and its
    walks must stay far inside their stack regions (5 words per claimed node)."""
    gt.graphtest_claim_device_stats.argtypes = [u32p, u32p]
    bases, seg_off, eso, _ = synth.poa_batch(seed, 2, depth=depth, length=length, err=err, length_jitter=0.1)
    for e in range(len(eso) - 1):
        run_host(gt, bases, seg_off[eso[e]: eso[e + 1] + 1])
    declined, peak = C.c_uint32(), C.c_uint32()
    gt.graphtest_claim_device_stats(C.byref(declined), C.byref(peak))
    assert declined.value == 0
    assert peak.value <= 400, f"a walk needed {peak.value / 100:.2f} stack words per claimed node"
