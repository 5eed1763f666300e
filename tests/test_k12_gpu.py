"""GPU parity of hgpu_compact_lr / hgpu_backbone_edges (through the C ABI) against the oracle and the reference text."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_io
import io_helpers
import oracle_ffi

pytestmark = pytest.mark.gpu


def cols(g, elems):
    return g["hits"]["t_id"][elems["hit"]], g["hits"]["is_rev"][elems["hit"]]


def test_compact_lr_golden(ctx, oracle):
    g = golden_io.inputs()
    got, goff = ctx.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"])
    ref, roff = oracle.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"])
    assert np.array_equal(goff, roff) and got.tobytes() == ref.tobytes()
    assert io_helpers.format_compact(got, goff, g["hits"]) == golden_io.text("syn200k_compact_uniq.txt")


@pytest.mark.parametrize("blk,sim,dev", [(100, 0.85, 0.15), (300, 0.90, 0.05), (900, 0.80, 0.5)])
def test_compact_lr_other_thresholds(ctx, oracle, blk, sim, dev):
    g = golden_io.inputs()
    kw = dict(min_aln_block=blk, min_aln_sim=sim, max_uniq_dev=dev)
    got, goff = ctx.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"], **kw)
    ref, roff = oracle.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"], **kw)
    assert np.array_equal(goff, roff) and got.tobytes() == ref.tobytes()


def test_backbone_edges_golden(ctx, oracle):
    g = golden_io.inputs()
    elems, off = ctx.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"])
    tid, rev = cols(g, elems)
    key, soff, supp, keep = ctx.backbone_edges(tid, rev, off, 3)
    rkey, rsoff, rsupp, rkeep = oracle.backbone_edges(tid, rev, off, 3)
    assert np.array_equal(key, rkey) and np.array_equal(soff, rsoff) and np.array_equal(keep, rkeep)
    assert supp.tobytes() == rsupp.tobytes()
    assert io_helpers.format_gfa_links(key) == golden_io.text("syn200k_backbone01.links")
    assert io_helpers.format_gfa_links(key[keep == 1]) == golden_io.text("syn200k_backbone02.links")


def test_backbone_edges_random_with_self_loops_and_repeats(ctx, oracle):
    """Random compact reads over few contigs: many supports per key, both strands, self loops (quirk Q6), empty reads."""
    rng = np.random.default_rng(5)
    # (2, 3000): ~1,500 supports per key - lists longer than 64 take the bitonic network of csrc/sort.cuh instead of the rank sort
    for n_contigs, n_reads in ((6, 400), (50, 2000), (3000, 5000), (2, 3000)):
        lens = rng.integers(0, 9, n_reads)
        off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint32)
        tid = rng.integers(0, n_contigs, int(off[-1])).astype(np.uint32)
        rev = rng.integers(0, 2, int(off[-1])).astype(np.uint8)
        for sup in (1, 3):
            key, soff, supp, keep = ctx.backbone_edges(tid, rev, off, sup)
            rkey, rsoff, rsupp, rkeep = oracle.backbone_edges(tid, rev, off, sup)
            assert np.array_equal(key, rkey) and np.array_equal(soff, rsoff) and np.array_equal(keep, rkeep)
            assert supp.tobytes() == rsupp.tobytes()


def test_backbone_edges_empty(ctx):
    key, soff, supp, keep = ctx.backbone_edges(np.zeros(0, np.uint32), np.zeros(0, np.uint8), np.zeros(1, np.uint32))
    assert len(key) == 0 and len(supp) == 0
    key, soff, supp, keep = ctx.backbone_edges(np.array([3], np.uint32), np.array([0], np.uint8), np.array([0, 0, 1, 1], np.uint32))
    assert len(key) == 0


def test_compact_lr_empty_and_ragged(ctx, oracle):
    g = golden_io.inputs()
    # keep only every third read's hits: most reads become empty or single-hit groups
    h, ro = g["hits"], g["read_off"]
    keep_rows = np.concatenate([np.arange(ro[r], ro[r + 1]) for r in range(0, g["n_reads"], 3)]).astype(np.int64)
    sub = {k: np.ascontiguousarray(v[keep_rows]) for k, v in h.items() if k not in ("cg_off", "cg_ops")}
    ops = [h["cg_ops"][h["cg_off"][i]: h["cg_off"][i + 1]] for i in keep_rows]
    sub["cg_off"] = np.concatenate(([0], np.cumsum([len(o) for o in ops]))).astype(np.uint32)
    sub["cg_ops"] = np.concatenate(ops).astype(np.uint32)
    cnt = np.zeros(g["n_reads"], dtype=np.int64)
    for r in range(0, g["n_reads"], 3):
        cnt[r] = ro[r + 1] - ro[r]
    sro = np.concatenate(([0], np.cumsum(cnt))).astype(np.uint32)
    got, goff = ctx.compact_lr(sub, sro, g["mean_kmer"], g["uniq_freq"])
    ref, roff = oracle.compact_lr(sub, sro, g["mean_kmer"], g["uniq_freq"])
    assert np.array_equal(goff, roff) and got.tobytes() == ref.tobytes()
    # no reads at all
    empty = {k: np.zeros(0, v.dtype) for k, v in h.items()}
    empty["cg_off"] = np.zeros(1, np.uint32); empty["cg_ops"] = np.zeros(1, np.uint32)
    got, goff = ctx.compact_lr(empty, np.zeros(1, np.uint32), g["mean_kmer"], g["uniq_freq"])
    assert len(got) == 0 and goff.tolist() == [0]


def test_compact_lr_and_backbone_fresh_synthetic(ctx, oracle):
    """A second seeded dataset generated on the box by the travelling generator binary (2 Mb, 6k reads)."""
    if not os.path.exists(oracle_ffi.GEN_BIN):
        pytest.skip("oracle/_ref/gen_synth not built")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([oracle_ffi.GEN_BIN, tmp, "2000000", "6000", "8000", "11"], check=True, stdout=subprocess.DEVNULL)
        lens, km, _, _ = io_helpers.load_contigs(os.path.join(tmp, "contigs.fa"))
        hits, ro = io_helpers.parse_paf(os.path.join(tmp, "map.paf"), 6000)
    uf = io_helpers.calc_uniq_freq(lens, km)
    got, goff = ctx.compact_lr(hits, ro, km, uf)
    ref, roff = oracle.compact_lr(hits, ro, km, uf)
    assert np.array_equal(goff, roff) and got.tobytes() == ref.tobytes()
    tid, rev = hits["t_id"][got["hit"]], hits["is_rev"][got["hit"]]
    key, soff, supp, keep = ctx.backbone_edges(tid, rev, goff, 3)
    rkey, rsoff, rsupp, rkeep = oracle.backbone_edges(tid, rev, goff, 3)
    assert np.array_equal(key, rkey) and np.array_equal(soff, rsoff) and np.array_equal(keep, rkeep) and supp.tobytes() == rsupp.tobytes()


@pytest.mark.parametrize("seed", [1, 2])
def test_adversarial_hits_text_to_edge_table(ctx, oracle, seed):
    """tests/golden/k1adv_*: PAF text -> hit table -> compact reads -> edge table, all on the GPU, against the reference
    binary's compact_uniq.txt and GFA links (heavy overlaps, sort ties, repeated contigs, rows at the filter thresholds)."""
    a = golden_io.k1_adversarial(seed)
    hits = ctx.parse_paf(a["paf"])
    read_off = np.searchsorted(hits["q_id"], np.arange(a["n_reads"] + 1), side="left").astype(np.uint32)
    uf = io_helpers.calc_uniq_freq(a["contig_len"], a["mean_kmer"])
    elems, off = ctx.compact_lr(hits, read_off, a["mean_kmer"], uf)
    assert io_helpers.format_compact(elems, off, hits) == a["compact"]
    key, soff, supp, keep = ctx.backbone_edges(hits["t_id"][elems["hit"]], hits["is_rev"][elems["hit"]], off, 3)
    assert io_helpers.format_gfa_links(key) == a["links01"]
    assert io_helpers.format_gfa_links(key[keep == 1]) == a["links02"]
