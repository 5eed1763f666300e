"""GPU parity of hgpu_edge_coords (through the C ABI) against the oracle, which is pinned on the reference's
log_coordinate.txt (tests/test_oracle_golden.py). Bit-exact: integer work."""
import numpy as np
import pytest

import coords_cases
import golden_io

pytestmark = pytest.mark.gpu


def check(ctx, oracle, c, hits):
    ref = oracle.edge_coords(c["edge_rev"], c["supp_off"], c["supp"], c["elems"], c["cl_off"], c["read_len"], hits)
    got = ctx.edge_coords(c["edge_rev"], c["supp_off"], c["supp"], c["elems"], c["cl_off"], c["read_len"], hits)
    assert got[0].tobytes() == ref[0].tobytes(), "per-edge intervals / anchor positions / counts differ"
    bad = np.nonzero(got[1] != ref[1])[0]
    assert len(bad) == 0, f"{len(bad)} supports differ, first {bad[:5]}"
    return ref


def test_edge_coords_golden(ctx, oracle):
    """The 120 edges the reference processed on the golden dataset, with the compact reads the GPU path itself produced."""
    ci = golden_io.coord_inputs(oracle)
    g = ci["g"]
    elems, off = ctx.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"])
    assert elems.tobytes() == ci["elems"].tobytes()
    ci = dict(ci, elems=elems, cl_off=off)
    ref = check(ctx, oracle, ci, g["hits"])
    gold = ci["gold"]
    assert [(int(e["c1"]), int(e["c2"]), int(e["n_best"])) for e in ref[0]] == [(x["c1"], x["c2"], x["n_best"]) for x in gold]


@pytest.mark.parametrize("seed,n_edges,max_supp", [(1, 200, 90), (2, 3000, 40), (3, 50, 700), (4, 1, 2), (5, 4, 3000)])
def test_edge_coords_stress(ctx, oracle, seed, n_edges, max_supp):
    """Ties, equal-depth optima, several bitmask words per edge, degenerate elements, refused walks, rows without CIGAR."""
    c = coords_cases.random_case(seed, n_edges=n_edges, max_supp=max_supp)
    check(ctx, oracle, c, c["hits"])


def test_edge_coords_empty_and_invalid(ctx):
    import haslr_b200
    z32 = np.zeros(1, np.uint32)
    oe, os_ = ctx.edge_coords(np.zeros(0, np.uint8), z32, np.zeros(0, haslr_b200.ffi.EDGE_SUPP), np.zeros(0, haslr_b200.ffi.CL_ELEM), z32, np.zeros(0, np.uint32),
                              dict(is_rev=np.zeros(0, np.uint8), cg_off=z32, cg_ops=z32))
    assert len(oe) == 0 and len(os_) == 0
    c = coords_cases.random_case(9, n_edges=4, max_supp=5)
    bad = c["supp"].copy(); bad["cmp_head"][0] = 1000       # element index outside its compact read: refused, not read
    with pytest.raises(haslr_b200.HgpuError) as ei:
        ctx.edge_coords(c["edge_rev"], c["supp_off"], bad, c["elems"], c["cl_off"], c["read_len"], c["hits"])
    assert ei.value.code == -1
