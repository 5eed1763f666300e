"""GPU parity of the batched POA path (hgpu_poa_batch, through the C ABI) against the CPU oracle. Bit-exact."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def check_batch(ctx, oracle, bases, seg_off, eso, scores=(5, -4, -8)):
    cons, off, status = ctx.poa_batch(bases, seg_off, eso, *scores)
    assert (status == 0).all(), f"non-zero edge status: {np.unique(status)}"
    rc, roff, cells, _ = oracle.poa_batch(bases, seg_off, eso, *scores, threads=8)
    bad = [e for e in range(len(eso) - 1)
           if cons[int(off[e]): int(off[e + 1])].tobytes() != rc[int(roff[e]): int(roff[e + 1])].tobytes()]
    assert not bad, f"{len(bad)} of {len(eso) - 1} consensus strings differ from the oracle, first: edge {bad[0]}"
    assert np.array_equal(off, roff)
    st = ctx.poa_stats()
    assert st["cells"] == cells
    return st


@pytest.mark.parametrize("n_prior", [1, 2, 5])
@pytest.mark.parametrize("force_i32", [0, 1, 2])
def test_score_matrix_and_alignment_match_oracle(ctx, oracle, n_prior, force_i32):
    """The whole DP matrix (in the reference's H space), the alignment and the graph it is computed on, in each of the
    three cell encodings (0 = picked: plain int16 here, 1 = int32, 2 = row-relative int16)."""
    bases, seg_off, eso, _ = synth.poa_batch(7, 1, depth=7, length=700, length_jitter=0.0)
    ref = oracle.poa_debug(bases, seg_off, n_prior)
    got = ctx.poa_debug(bases, seg_off, n_prior, force_i32=force_i32)
    assert got["V"] == ref["V"] and got["L"] == ref["L"]
    for k in ("rank2node", "code", "pred_off", "pred_node", "pred_weight"):
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["H"], ref["H"])
    assert np.array_equal(got["aln_node"], ref["aln_node"])
    assert np.array_equal(got["aln_pos"], ref["aln_pos"])


def test_score_matrix_multi_stripe(ctx, oracle):
    """L > 512 columns: several stripes, boundary-column hand-off between them."""
    bases, seg_off, eso, _ = synth.poa_batch(11, 1, depth=4, length=1500)
    for force in (0, 1, 2):
        ref = oracle.poa_debug(bases, seg_off, 3)
        got = ctx.poa_debug(bases, seg_off, 3, force_i32=force)
        assert np.array_equal(got["H"], ref["H"])
        assert np.array_equal(got["aln_node"], ref["aln_node"]) and np.array_equal(got["aln_pos"], ref["aln_pos"])


def test_consensus_small_batch(ctx, oracle):
    bases, seg_off, eso, _ = synth.poa_batch(3, 64, depth=6, length=400, length_jitter=0.3)
    check_batch(ctx, oracle, bases, seg_off, eso)


def test_consensus_cfg3_shape(ctx, oracle):
    """BASELINE config 3 shape (6 supporting reads x 1.5 kb), a slice the oracle finishes in seconds."""
    bases, seg_off, eso, _ = synth.poa_batch(5, 96, depth=6, length=1500)
    st = check_batch(ctx, oracle, bases, seg_off, eso)
    assert st["alignments"] == 96 * 5 and st["alignments_i32"] == 0


def test_consensus_high_error_and_depth(ctx, oracle):
    bases, seg_off, eso, _ = synth.poa_batch(9, 48, depth=24, length=300, err=(0.08, 0.06, 0.04), length_jitter=0.3, depth_jitter=6)
    check_batch(ctx, oracle, bases, seg_off, eso)


def test_consensus_wide_range(ctx, oracle):
    """Segments long enough that the plain int16 score range does not hold (SPOA switches to int32 lanes there): these
    alignments run in row-relative int16 cells, and in int32 cells when that is forced."""
    bases, seg_off, eso, _ = synth.poa_batch(13, 3, depth=3, length=4000)
    st = check_batch(ctx, oracle, bases, seg_off, eso)
    assert st["alignments_rel16"] > 0 and st["alignments_i32"] == 0


def test_score_matrix_wide_range_rel16(ctx, oracle):
    """A matrix whose range needs the row-relative encoding (13 (L+1) + 8 (V+2) > 64000), cell by cell."""
    bases, seg_off, eso, _ = synth.poa_batch(17, 1, depth=4, length=3300)
    ref = oracle.poa_debug(bases, seg_off, 3)
    got = ctx.poa_debug(bases, seg_off, 3)
    assert np.array_equal(got["H"], ref["H"])
    assert np.array_equal(got["aln_node"], ref["aln_node"]) and np.array_equal(got["aln_pos"], ref["aln_pos"])


def test_consensus_deep_long_edges_rel16(ctx, oracle):
    """BASELINE config 2 shape: 25x coverage over 2-3 kb gaps, graphs of several thousand nodes."""
    bases, seg_off, eso, _ = synth.poa_batch(19, 6, depth=14, length=2600, length_jitter=0.2)
    st = check_batch(ctx, oracle, bases, seg_off, eso)
    assert st["alignments_rel16"] > 0


@pytest.mark.parametrize("mode", [1, 2])
def test_forced_encodings_whole_batches(oracle, monkeypatch, mode):
    """Every alignment of a batch forced into int32 (1) / row-relative int16 (2) cells: same consensus as the oracle."""
    import haslr_b200
    monkeypatch.setenv("HGPU_FORCE_MODE", str(mode))
    c = haslr_b200.Context(0)
    try:
        b, so, eo, _ = synth.poa_batch(3, 64, depth=6, length=400, length_jitter=0.3)
        st = check_batch(c, oracle, b, so, eo)
        assert st["alignments_rel16" if mode == 2 else "alignments_i32"] == st["alignments"]
        b, so, eo, _ = synth.poa_batch(9, 32, depth=24, length=300, err=(0.08, 0.06, 0.04), length_jitter=0.3, depth_jitter=6)
        check_batch(c, oracle, b, so, eo)
        b, so, eo, _ = synth.poa_batch(5, 24, depth=6, length=1500)
        check_batch(c, oracle, b, so, eo)
        rng = np.random.default_rng(23)               # unrelated segments: many predecessor-free ranks and far predecessors
        b, so, eo = synth.from_strings([[bytes(synth.ACGT[rng.integers(0, 4, 700)]) for _ in range(6)] for _ in range(4)])
        check_batch(c, oracle, b, so, eo)
    finally:
        c.close()


def test_edge_cases(ctx, oracle):
    edges = [
        [],                                          # edge without supporting segments -> empty consensus
        [b""],                                       # only an empty segment
        [b"ACGT"],                                   # a single segment is its own consensus
        [b"A", b"A", b"A"],
        [b"ACGTACGT", b"", b"ACGTACGT"],             # empty segments are skipped (Assemble.cpp:537)
        [b"AAAAAAAAAA", b"TTTTTTTTTT", b"AAAAAAAAAA"],
        [b"ACGTNNACGT", b"ACGTAAACGT", b"acgtaaacgt"],
        [b"ACGT" * 300, b"ACGT" * 10, b"ACGT" * 500],  # ragged
        [b"GATTACA", b"GATACA", b"GATTTACA", b"CATTACA", b"GATTACAT", b"TGATTACA"],
    ]
    bases, seg_off, eso = synth.from_strings(edges)
    check_batch(ctx, oracle, bases, seg_off, eso)


def test_empty_batch(ctx):
    cons, off, status = ctx.poa_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64), np.zeros(1, np.uint32))
    assert len(cons) == 0 and off.tolist() == [0] and len(status) == 0


def test_bases_from_the_staging_buffer(ctx):
    """hgpu_host_staging: the context's page-locked buffers. Same result whether the bases come from pageable memory or from the
    staging buffer; asking for less returns the same buffer, asking for more grows it; a bad index is refused."""
    import ctypes
    import haslr_b200
    bases, seg_off, eso, _ = synth.poa_batch(29, 48, depth=6, length=500, length_jitter=0.3)
    cons, off, status = ctx.poa_batch(bases, seg_off, eso)
    p = ctx.host_staging(0, bases.nbytes)
    assert p and ctx.host_staging(0, bases.nbytes // 2) == p
    ctypes.memmove(p, bases.ctypes.data, bases.nbytes)
    pinned = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(bases.nbytes,))
    cons2, off2, status2 = ctx.poa_batch(pinned, seg_off, eso)
    assert np.array_equal(off, off2) and np.array_equal(status, status2) and cons.tobytes() == cons2.tobytes()
    q = ctx.host_staging(1, 1 << 20)
    assert q and q != p
    big = ctx.host_staging(0, bases.nbytes * 4)
    ctypes.memset(big, 0, bases.nbytes * 4)
    with pytest.raises(haslr_b200.HgpuError):
        ctx.host_staging(2, 16)


def test_other_scores(ctx, oracle):
    bases, seg_off, eso, _ = synth.poa_batch(17, 16, depth=5, length=300, length_jitter=0.2)
    for scores in ((3, -5, -4), (1, -1, -1), (2, -6, -2)):
        check_batch(ctx, oracle, bases, seg_off, eso, scores)


def test_band_is_rejected(ctx):
    import haslr_b200
    with pytest.raises(haslr_b200.HgpuError) as ei:
        ctx.poa_batch(np.frombuffer(b"ACGT", np.uint8), np.array([0, 4], np.uint64), np.array([0, 1], np.uint32), band=64)
    assert ei.value.code == -5


def test_growth_retry_path(ctx, oracle):
    """Unrelated segments make the node count grow far beyond the first estimate: the scheduler must retry, not fail."""
    rng = np.random.default_rng(23)
    edges = [[bytes(synth.ACGT[rng.integers(0, 4, 400)]) for _ in range(8)] for _ in range(4)]
    bases, seg_off, eso = synth.from_strings(edges)
    check_batch(ctx, oracle, bases, seg_off, eso)


def test_roundtrip_property_large(ctx):
    """Size-independent property at a size the oracle would take minutes for: consensus of R identical strings is
    that string, and every consensus of noisy copies is closer to the truth than a single read's expected error."""
    rng = np.random.default_rng(29)
    truths = [synth.ACGT[rng.integers(0, 4, 1200)] for _ in range(512)]
    edges = [[t.tobytes()] * 4 for t in truths]
    bases, seg_off, eso = synth.from_strings(edges)
    cons, off, status = ctx.poa_batch(bases, seg_off, eso)
    assert (status == 0).all()
    for e, t in enumerate(truths):
        assert cons[int(off[e]): int(off[e + 1])].tobytes() == t.tobytes()


@pytest.mark.parametrize("force_i32", [0, 1, 2])
def test_deep_kernel_score_matrix_and_batches(oracle, monkeypatch, force_i32):
    """k_poa_edges_deep (every row parked in a shared-memory ring; normally edges with >= 10 supporting reads), forced
    onto small inputs: score matrix and alignment cell by cell in each encoding, then whole batches."""
    import haslr_b200
    monkeypatch.setenv("HGPU_DEEP_MIN_READS", "2")
    if force_i32:
        monkeypatch.setenv("HGPU_FORCE_MODE", str(force_i32))
    c = haslr_b200.Context(0)
    try:
        bases, seg_off, eso, _ = synth.poa_batch(7, 1, depth=7, length=700)
        for n_prior in (1, 4, 6):
            ref = oracle.poa_debug(bases, seg_off, n_prior)
            got = c.poa_debug(bases, seg_off, n_prior, force_i32=force_i32)
            assert np.array_equal(got["H"], ref["H"])
            assert np.array_equal(got["aln_node"], ref["aln_node"]) and np.array_equal(got["aln_pos"], ref["aln_pos"])
        b, so, eo, _ = synth.poa_batch(9, 32, depth=24, length=300, err=(0.08, 0.06, 0.04), length_jitter=0.3, depth_jitter=6)
        check_batch(c, oracle, b, so, eo)
        b, so, eo, _ = synth.poa_batch(5, 24, depth=6, length=1500)
        check_batch(c, oracle, b, so, eo)
    finally:
        c.close()


def test_team_kernel_matches_oracle(oracle, monkeypatch):
    """Edges large enough for the block-per-edge kernel (forced here with a low threshold): multi-stripe score matrices
    filled by 8 warps in a pipeline must give the same consensus as the warp-per-edge kernel and the oracle."""
    import haslr_b200
    monkeypatch.setenv("HGPU_POOL", "0")           # the pool kernel takes these edges otherwise
    monkeypatch.setenv("HGPU_TEAM", "8")
    monkeypatch.setenv("HGPU_TEAM_MIN_CELLS", "1000")
    c = haslr_b200.Context(0)
    try:
        # mixed batch: long gaps (4-10 stripes, some int32) go to the team kernel, short ones stay warp-per-edge
        b1, so1, es1, _ = synth.poa_batch(31, 6, depth=5, length=2600, length_jitter=0.3)
        b2, so2, es2, _ = synth.poa_batch(32, 40, depth=6, length=300, length_jitter=0.3)
        b3, so3, es3, _ = synth.poa_batch(33, 2, depth=3, length=5200)
        bases = np.concatenate((b1, b2, b3))
        seg_off = np.concatenate((so1, so2[1:] + so1[-1], so3[1:] + so1[-1] + so2[-1])).astype(np.uint64)
        eso = np.concatenate((es1, es2[1:] + es1[-1], es3[1:] + es1[-1] + es2[-1])).astype(np.uint32)
        st = check_batch(c, oracle, bases, seg_off, eso)
        assert st["dp_launches"] >= 2          # team kernel + warp-per-edge kernel
        assert st["alignments_rel16"] > 0      # the long gaps run in row-relative int16 cells, stripes pipelined over the team
    finally:
        c.close()


@pytest.mark.parametrize("ctx_per_block", [0, 1, 3, 8])
def test_pool_kernel_matches_oracle(oracle, monkeypatch, ctx_per_block):
    """k_poa_pool (a block's warps pull stripe tasks and graph tasks from several edges in flight), forced onto every edge of mixed
    batches: single-stripe and 4-10-stripe alignments, edges without / with one segment, more edges than contexts and fewer."""
    import haslr_b200
    monkeypatch.setenv("HGPU_DEEP_MIN_READS", "0")
    if ctx_per_block:
        monkeypatch.setenv("HGPU_POOL_CTX", str(ctx_per_block))
    c = haslr_b200.Context(0)
    try:
        b1, so1, es1, _ = synth.poa_batch(31, 6, depth=5, length=2600, length_jitter=0.3)
        b2, so2, es2, _ = synth.poa_batch(32, 40, depth=6, length=300, length_jitter=0.3)
        b3, so3, es3, _ = synth.poa_batch(33, 2, depth=3, length=5200)
        bases = np.concatenate((b1, b2, b3))
        seg_off = np.concatenate((so1, so2[1:] + so1[-1], so3[1:] + so1[-1] + so2[-1])).astype(np.uint64)
        eso = np.concatenate((es1, es2[1:] + es1[-1], es3[1:] + es1[-1] + es2[-1])).astype(np.uint32)
        st = check_batch(c, oracle, bases, seg_off, eso)
        assert st["alignments_rel16"] == st["alignments"] > 0
        edges = [[], [b"ACGT"], [b""], [b"ACGTACGT", b"", b"ACGTACGT"], [b"ACGT" * 300, b"ACGT" * 10, b"ACGT" * 500],
                 [b"GATTACA", b"GATACA", b"GATTTACA", b"CATTACA", b"GATTACAT", b"TGATTACA"]]
        bases, seg_off, eso = synth.from_strings(edges)
        check_batch(c, oracle, bases, seg_off, eso)
        b, so, eo, _ = synth.poa_batch(9, 48, depth=24, length=300, err=(0.08, 0.06, 0.04), length_jitter=0.3, depth_jitter=6)
        check_batch(c, oracle, b, so, eo)
        b, so, eo, _ = synth.poa_batch(19, 6, depth=14, length=2600, length_jitter=0.2)
        check_batch(c, oracle, b, so, eo)
        check_batch(c, oracle, b, so, eo, (3, -5, -4))
    finally:
        c.close()
