"""CPU check of the product's per-edge coordinate core (haslr_b200/csrc/coords_core.cuh, __host__ __device__) against
the oracle (itself pinned on the reference's log_coordinate.txt): golden dataset and seeded stress cases."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import coords_cases
import golden_io
import oracle_ffi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "coords_host_check.cpp")
LIB = os.path.join(HERE, "native", "libcoordstest.so")
u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def k4():
    hdr = os.path.join(HERE, "..", "haslr_b200", "csrc", "coords_core.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", LIB, SRC], check=True)
    L = C.CDLL(LIB)
    L.coordshost_edge_coords.restype = C.c_int
    L.coordshost_edge_coords.argtypes = [C.c_uint32, u8p, u32p, C.c_void_p, C.c_void_p, u32p, u32p, u8p, u32p, u32p, C.c_void_p, C.c_void_p]
    return L


def run_host(L, c, hits):
    p = lambda a, t: a.ctypes.data_as(t)
    n = len(c["edge_rev"])
    oe = np.zeros(max(n, 1), dtype=oracle_ffi.EDGE_COORD); os_ = np.zeros(max(len(c["supp"]), 1), dtype=oracle_ffi.SUPP_COORD)
    arrs = [np.ascontiguousarray(c[k]) for k in ("edge_rev", "supp_off", "supp", "elems", "cl_off", "read_len")]
    L.coordshost_edge_coords(n, p(arrs[0], u8p), p(arrs[1], u32p), arrs[2].ctypes.data, arrs[3].ctypes.data, p(arrs[4], u32p), p(arrs[5], u32p),
                             p(hits["is_rev"], u8p), p(hits["cg_off"], u32p), p(hits["cg_ops"], u32p), oe.ctypes.data, os_.ctypes.data)
    return oe[:n], os_[: len(c["supp"])]


def same(a, b):
    return a.tobytes() == b.tobytes()


def test_core_matches_oracle_on_golden(k4, oracle):
    ci = golden_io.coord_inputs(oracle)
    hits = ci["g"]["hits"]
    ref = oracle.edge_coords(ci["edge_rev"], ci["supp_off"], ci["supp"], ci["elems"], ci["cl_off"], ci["read_len"], hits)
    got = run_host(k4, ci, hits)
    assert same(got[0], ref[0]) and same(got[1], ref[1])


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_core_matches_oracle_on_stress_cases(k4, oracle, seed):
    c = coords_cases.random_case(seed)
    ref = oracle.edge_coords(c["edge_rev"], c["supp_off"], c["supp"], c["elems"], c["cl_off"], c["read_len"], c["hits"])
    got = run_host(k4, c, c["hits"])
    assert same(got[0], ref[0])
    bad = np.nonzero(got[1] != ref[1])[0]
    assert len(bad) == 0, f"{len(bad)} supports differ, first {bad[:5]}: {got[1][bad[:3]]} vs {ref[1][bad[:3]]}"
    assert (ref[1]["in_best"] == 1).sum() > 100 and (ref[0]["n_cns"] < ref[0]["n_best"]).any()       # accepted and refused walks both occur
