"""GPU parity of hgpu_paf_tokenize / hgpu_paf_fetch (through the C ABI) against the oracle (reference-style getline / split /
str2type), which reproduces the hit table the golden fixtures were made from (tests/test_paf_host.py). Bit-exact."""
import numpy as np
import pytest

import golden_io
import paf_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold_text():
    return paf_cases.golden_paf()


def test_tokeniser_golden_and_edge_cases(ctx, oracle, gold_text):
    g = golden_io.inputs()
    got = ctx.parse_paf(gold_text)
    for k in ("q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block", "is_rev", "mapq", "cg_off"):
        assert np.array_equal(got[k], g["hits"][k]), k
    # text that ends without a line feed, ends inside the first 256-byte chunk, has empty lines, odd tags ...
    for text in (gold_text, gold_text[:-1], paf_cases.ODD, paf_cases.ODD + b"\n", b"", b"\n\n", b"\n" + paf_cases.ODD,
                 gold_text[:50000].rsplit(b"\n", 1)[0], gold_text[:255], gold_text[:256].rsplit(b"\t", 1)[0]):
        ref, n = oracle.parse_paf(text)
        if ref is None:                         # a cut that leaves fewer than 12 columns: both refuse
            import haslr_b200
            with pytest.raises(haslr_b200.HgpuError):
                ctx.parse_paf(text)
            continue
        got = ctx.parse_paf(text)
        assert len(got["q_id"]) == n and paf_cases.same_hits(got, ref)


def test_tokeniser_feeds_compact_lr(ctx, oracle, gold_text):
    """text -> hit table -> compact long reads on the GPU equals the reference's compact_uniq.txt."""
    import io_helpers
    g = golden_io.inputs()
    hits = ctx.parse_paf(gold_text)
    read_off = np.searchsorted(hits["q_id"], np.arange(g["n_reads"] + 1), side="left").astype(np.uint32)
    elems, off = ctx.compact_lr(hits, read_off, g["mean_kmer"], g["uniq_freq"])
    assert io_helpers.format_compact(elems, off, hits) == golden_io.text("syn200k_compact_uniq.txt")


def test_tokeniser_large_random(ctx, oracle):
    """A few MB of generated rows with long CIGARs: many chunks, lines spanning chunk borders."""
    rng = np.random.default_rng(3)
    rows = []
    for i in range(20000):
        ops = "".join("%d%s" % (rng.integers(1, 300), "MID"[j % 3 if j % 2 else 0]) for j in range(int(rng.integers(1, 60))))
        rows.append("%d\t%d\t%d\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\ttp:A:P\tcg:Z:%s" % (
            i // 7, rng.integers(1000, 20000), rng.integers(0, 500), rng.integers(500, 9000), "+-"[i & 1], rng.integers(0, 5000),
            rng.integers(100, 3000), rng.integers(0, 50), rng.integers(50, 3000), rng.integers(0, 3000), rng.integers(1, 3000), rng.integers(0, 61), ops))
    text = ("\n".join(rows) + "\n").encode()
    ref, n = oracle.parse_paf(text)
    got = ctx.parse_paf(text)
    assert n == 20000 and paf_cases.same_hits(got, ref)


def test_tokeniser_fuzzed_lines(ctx, oracle):
    """Rows assembled from odd tokens (signs, blanks, overflow, empty columns, carriage returns, tags in any order)."""
    text, _ = paf_cases.fuzzed(11)
    ref, n = oracle.parse_paf(text)
    got = ctx.parse_paf(text)
    assert n == 4000 and len(got["q_id"]) == n and paf_cases.same_hits(got, ref)


def test_short_line_is_refused(ctx):
    import haslr_b200
    with pytest.raises(haslr_b200.HgpuError) as ei:
        ctx.parse_paf(b"1\t2\t3\t4\t+\t5\t6\t7\t8\t9\t10\t11\tcg:Z:5M\n1\t2\t3\n")
    assert ei.value.code == -1 and "line 2" in str(ei.value)


def test_cigar_run_total_beyond_32_bits_is_refused(monkeypatch):
    """cg_off is 32-bit: a buffer whose CIGAR runs do not fit must fail with HGPU_E_UNSUPPORTED, not wrap (the totals are
    scanned in 64 bits; the test adds a bias to the real total instead of tokenising 16 GB of text)."""
    import haslr_b200
    monkeypatch.setenv("HGPU_TEST_PAF_OPS_BIAS", str(2**32 - 2))
    c = haslr_b200.Context(0)
    try:
        line = b"0\t100\t0\t50\t+\t1\t200\t0\t50\t50\t50\t60\tcg:Z:20M1I29M\n"
        with pytest.raises(haslr_b200.HgpuError) as ei:
            c.tokenize(line * 3)
        assert ei.value.code == -5
    finally:
        c.close()


def test_rows_out_of_read_order_are_refused(ctx):
    import haslr_b200
    row = lambda q: b"%d\t100\t0\t50\t+\t1\t200\t0\t50\t50\t50\t60\tcg:Z:50M\n" % q
    ctx.tokenize(row(0) + row(2) + row(1))
    with pytest.raises(haslr_b200.HgpuError) as ei:
        ctx.hits_group(3)
    assert ei.value.code == -1
    ctx.tokenize(row(0) + row(2) + row(2))
    assert ctx.hits_group(4).tolist() == [0, 1, 1, 3, 3]
    with pytest.raises(haslr_b200.HgpuError):
        ctx.hits_group(2)                    # read 2 named, two reads loaded
