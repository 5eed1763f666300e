"""The CUDA kernels (through the C ABI) against what the REFERENCE binary logged, not against the oracle:

* tests/golden/syn200k_poa.txt.gz — the 120 segment groups (real depth spread, 3-40 reads) the reference fed to its POA
  calls and the consensus it wrote to log_consensus.txt (reference linked with the restated SPOA: K3 stays "parity partial");
* tests/golden/k4adv_* — log_coordinate.txt of the reference on 20-40 reads per edge with tied alignment ends and
  equal-depth optima (where the `>=` of the head sweep and the `>` of the tail sweep differ, Assemble.cpp:45 vs :97).

And the SURVEY §8(c) invariants that need no SPOA at all (synthetic truth known): the consensus is at least as close to
the truth as any single read, and its length stays within a few bases of the truth's.
"""
import ctypes as C

import numpy as np
import pytest

import golden_io
import io_helpers
import oracle_ffi
import synth
from test_asm_host import asm, u32p, u64p          # noqa: F401  (fixture: the product's host graph code)

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------------------
# K3 on the reference's own call sequence
# ---------------------------------------------------------------------------------------------------------------------
def _golden_poa(c):
    edges = golden_io.poa_edges()
    assert len(edges) >= 120
    bases, seg_off, eso = synth.from_strings([e[1] for e in edges])
    cons, off, status = c.poa_batch(bases, seg_off, eso)
    assert (status == 0).all()
    bad = [e[0] for i, e in enumerate(edges) if cons[int(off[i]): int(off[i + 1])].tobytes() != e[2]]
    assert not bad, f"{len(bad)} of {len(edges)} consensus strings differ from the reference's log_consensus.txt, first: {bad[0]}"
    return c.poa_stats()


def test_poa_golden_groups_default_kernels(ctx):
    st = _golden_poa(ctx)
    assert st["alignments"] > 1000


@pytest.mark.parametrize("env", [{"HGPU_DEEP_MIN_READS": "2"},                                       # every edge in the deep kernel
                                 {"HGPU_DEEP_MIN_READS": "2", "HGPU_FORCE_MODE": "2"},              # ... in row-relative int16 cells
                                 {"HGPU_POOL": "0", "HGPU_TEAM": "8", "HGPU_TEAM_MIN_CELLS": "1000", "HGPU_TEAM_ALPHA": "0.01"},
                                 {"HGPU_POOL": "0", "HGPU_TEAM": "4", "HGPU_TEAMS_PER_SM": "4", "HGPU_TEAM_MIN_CELLS": "1000", "HGPU_TEAM_ALPHA": "0.01"},
                                 {"HGPU_FORCE_MODE": "1"}])                                         # int32 cells
def test_poa_golden_groups_every_kernel(monkeypatch, env):
    import haslr_b200
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c = haslr_b200.Context(0)
    try:
        _golden_poa(c)
    finally:
        c.close()


# ---------------------------------------------------------------------------------------------------------------------
# K0 -> K1 -> K2 -> (host graph) -> K4 on the adversarial coordinate fixtures, against the reference's log
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [1, 2])
def test_k4_adversarial_on_device_matches_reference_log(ctx, asm, seed, tmp_path):
    a = golden_io.k4_adversarial(seed)
    hits = ctx.parse_paf(a["paf"])
    n_reads = len(a["read_len"])
    read_off = np.searchsorted(hits["q_id"], np.arange(n_reads + 1), side="left").astype(np.uint32)
    elems, off = ctx.compact_lr(hits, read_off, a["mean_kmer"], io_helpers.calc_uniq_freq(a["contig_len"], a["mean_kmer"]))
    key, soff, supp, _ = ctx.backbone_edges(hits["t_id"][elems["hit"]], hits["is_rev"][elems["hit"]], off, 3)
    p = lambda x, t: x.ctypes.data_as(t)
    coff = np.concatenate(([0], np.cumsum(a["contig_len"].astype(np.uint64)))).astype(np.uint64)
    asm.asmhost_prepare.argtypes = [C.c_uint32, C.c_char_p, u64p, C.c_uint64, u64p, u32p, C.c_void_p, C.c_uint32, C.c_char_p]
    n = asm.asmhost_prepare(len(a["contig_len"]), b"A" * int(coff[-1]), p(coff, u64p), len(key), p(key, u64p), p(soff, u32p), supp.ctypes.data, 3,
                            str(tmp_path).encode())
    gold = a["gold"]
    assert n == len(gold) and n >= 30
    e4 = np.zeros(4 * n, dtype=np.uint32); eso = np.zeros(n + 1, dtype=np.uint32); esupp = np.zeros(len(supp), dtype=oracle_ffi.EDGE_SUPP)
    asm.asmhost_edges.argtypes = [u32p, u32p, C.c_void_p, C.c_uint32]
    ns = asm.asmhost_edges(p(e4, u32p), p(eso, u32p), esupp.ctypes.data, len(esupp))
    e4 = e4.reshape(n, 4); esupp = esupp[:ns]
    assert [tuple(r) for r in e4.tolist()] == [x["edge"] for x in gold]
    oe, os_ = ctx.edge_coords((e4[:, 1] | (e4[:, 3] << 1)).astype(np.uint8), eso, esupp, elems, off, a["read_len"], hits)
    for e, ge in enumerate(gold):
        b, m = int(eso[e]), int(eso[e + 1] - eso[e])
        assert m == ge["n_supp"]
        assert ((oe[e]["int1_lo"], oe[e]["int1_hi"]), (oe[e]["int2_lo"], oe[e]["int2_hi"])) == (ge["int1"], ge["int2"]), e
        assert (oe[e]["c1"], oe[e]["c2"], oe[e]["n_best"]) == (ge["c1"], ge["c2"], ge["n_best"]), e
        best = []
        for i in range(m):
            o = os_[b + i]
            if o["in_best"]:
                rid = int(esupp[b + i]["lr_id_strand"]) & 0x7FFFFFFF
                ok = o["lr_start"] != -1 and o["lr_end"] != -1
                best.append((rid, int(a["read_len"][rid]), int(o["lr_strand"])) + ((int(o["lr_start"]) + 1, int(o["lr_end"]) - 1) if ok else (None, None)))
        assert best == ge["best"], e


@pytest.mark.parametrize("seed", [1, 2])
def test_resident_stages_equal_host_pointer_stages(asm, seed, tmp_path):
    """The device-resident chain the binary runs (hgpu_paf_tokenize -> hgpu_hits_group -> hgpu_compact_lr_dev ->
    hgpu_backbone_edges_dev -> hgpu_edge_coords_dev: nothing of the hit table comes back to the host) against the host-pointer
    entry points on the same adversarial PAF text: identical offsets, elements, edge table and coordinates."""
    import haslr_b200
    a = golden_io.k4_adversarial(seed)
    n_reads = len(a["read_len"])
    uf = io_helpers.calc_uniq_freq(a["contig_len"], a["mean_kmer"])
    c1, c2 = haslr_b200.Context(0), haslr_b200.Context(0)
    try:
        hits = c1.parse_paf(a["paf"])
        read_off = np.searchsorted(hits["q_id"], np.arange(n_reads + 1), side="left").astype(np.uint32)
        elems, off = c1.compact_lr(hits, read_off, a["mean_kmer"], uf)
        key, soff, supp, keep = c1.backbone_edges(hits["t_id"][elems["hit"]], hits["is_rev"][elems["hit"]], off, 3)
        rows, ops = c2.tokenize(a["paf"])
        assert rows == len(hits["q_id"]) and ops == int(hits["cg_off"][-1])
        assert np.array_equal(c2.hits_group(n_reads), read_off)
        delems, dtid, drev, doff = c2.compact_lr_dev(n_reads, rows, a["mean_kmer"], uf)
        assert np.array_equal(doff, off) and delems.tobytes() == elems.tobytes()
        assert np.array_equal(dtid, hits["t_id"][elems["hit"]]) and np.array_equal(drev, hits["is_rev"][elems["hit"]])
        dkey, dsoff, dsupp, dkeep = c2.backbone_edges_dev(doff, 3)
        assert np.array_equal(dkey, key) and np.array_equal(dsoff, soff) and np.array_equal(dkeep, keep) and dsupp.tobytes() == supp.tobytes()
        p = lambda x, t: x.ctypes.data_as(t)
        coff = np.concatenate(([0], np.cumsum(a["contig_len"].astype(np.uint64)))).astype(np.uint64)
        asm.asmhost_prepare.argtypes = [C.c_uint32, C.c_char_p, u64p, C.c_uint64, u64p, u32p, C.c_void_p, C.c_uint32, C.c_char_p]
        n = asm.asmhost_prepare(len(a["contig_len"]), b"A" * int(coff[-1]), p(coff, u64p), len(key), p(key, u64p), p(soff, u32p), supp.ctypes.data, 3,
                                str(tmp_path).encode())
        e4 = np.zeros(4 * n, dtype=np.uint32); eso = np.zeros(n + 1, dtype=np.uint32); esupp = np.zeros(len(supp), dtype=oracle_ffi.EDGE_SUPP)
        asm.asmhost_edges.argtypes = [u32p, u32p, C.c_void_p, C.c_uint32]
        ns = asm.asmhost_edges(p(e4, u32p), p(eso, u32p), esupp.ctypes.data, len(esupp))
        e4 = e4.reshape(n, 4); esupp = esupp[:ns]
        rev = (e4[:, 1] | (e4[:, 3] << 1)).astype(np.uint8)
        oe, os_ = c1.edge_coords(rev, eso, esupp, elems, off, a["read_len"], hits)
        doe, dos = c2.edge_coords_dev(rev, eso, esupp, a["read_len"])
        assert doe.tobytes() == oe.tobytes() and dos.tobytes() == os_.tobytes()
        st = c2.stage_stats()
        assert st["k0_rows"] == rows and st["k1_hits"] == rows and st["k4_supports"] == ns
        # a support that names an element outside its compact read is refused by the kernel, not read
        bad = esupp.copy(); bad["cmp_head"][0] = 1000
        with pytest.raises(haslr_b200.HgpuError) as ei:
            c2.edge_coords_dev(rev, eso, bad, a["read_len"])
        assert ei.value.code == -1
    finally:
        c1.close(); c2.close()


# ---------------------------------------------------------------------------------------------------------------------
# invariants against the synthetic truth (no SPOA, no oracle involved)
# ---------------------------------------------------------------------------------------------------------------------
def edit_distance(a, b):
    """Unit-cost Levenshtein distance of two uint8 arrays, one numpy row per base of `a` (the insertion chain of a row is
    a running minimum of D - j)."""
    n = len(b)
    j = np.arange(n + 1, dtype=np.int64)
    prev = j.copy()
    for i in range(1, len(a) + 1):
        t = np.empty(n + 1, dtype=np.int64)
        t[0] = i
        np.minimum(prev[1:] + 1, prev[:-1] + (b != a[i - 1]), out=t[1:])
        prev = np.minimum.accumulate(t - j) + j
    return int(prev[n])


@pytest.mark.parametrize("shape", [dict(n=24, depth=6, length=1500, err=(0.04, 0.03, 0.02)),        # BASELINE config 3
                                   dict(n=12, depth=6, length=1500, err=(0.05, 0.04, 0.03)),        # its 12 %-error variant
                                   dict(n=4, depth=28, length=2500, err=(0.04, 0.03, 0.02))])       # config 2's median edge
def test_consensus_is_closer_to_the_truth_than_any_read(ctx, shape):
    bases, seg_off, eso, truths = synth.poa_batch(41, shape["n"], depth=shape["depth"], length=shape["length"], err=shape["err"])
    cons, off, status = ctx.poa_batch(bases, seg_off, eso)
    assert (status == 0).all()
    for e, truth in enumerate(truths):
        c = cons[int(off[e]): int(off[e + 1])]
        reads = [bases[int(seg_off[s]): int(seg_off[s + 1])] for s in range(int(eso[e]), int(eso[e + 1]))]
        d_reads = [edit_distance(r, truth) for r in reads[:8]]          # eight reads bound the cost at depth 28
        d_cons = edit_distance(c, truth)
        assert d_cons <= min(d_reads), (e, d_cons, d_reads)
        assert d_cons <= 0.6 * float(np.median(d_reads)), (e, d_cons, d_reads)   # POA halves the error at depth 6, far more at 28
        med = float(np.median([len(r) for r in reads]))
        assert abs(len(c) - med) <= 0.03 * med + 8, (e, len(c), med)
        assert abs(len(c) - len(truth)) <= max(8, d_cons), (e, len(c), len(truth))
