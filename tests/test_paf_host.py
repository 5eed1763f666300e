"""CPU checks of the PAF tokeniser: the oracle (reference-style getline / split / istringstream) against the parse the golden
fixtures were made from, and the product's per-line core (haslr_b200/csrc/paf_core.cuh, __host__ __device__) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_io
import oracle_ffi
import paf_cases

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "paf_host_check.cpp")
LIB = os.path.join(HERE, "native", "libpaftest.so")


@pytest.fixture(scope="module")
def pafhost():
    hdr = os.path.join(HERE, "..", "haslr_b200", "csrc", "paf_core.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", LIB, SRC], check=True)
    return C.CDLL(LIB)


@pytest.fixture(scope="module")
def gold_text():
    if not os.path.exists(oracle_ffi.GEN_BIN):
        oracle_ffi.build(("tools",))
    return paf_cases.golden_paf()


def test_oracle_reproduces_the_golden_hit_table(oracle, gold_text):
    """The hit table the K1/K2/K4 goldens were made from (and through them the reference's compact_uniq.txt) is what the
    oracle reads out of the same map.paf."""
    g = golden_io.inputs()
    h, n = oracle.parse_paf(gold_text)
    assert n == len(g["hits"]["q_start"]) == 9912
    for k in ("q_start", "q_end", "t_id", "t_len", "t_start", "t_end", "n_match", "n_block", "is_rev", "mapq", "cg_off"):
        assert np.array_equal(h[k], g["hits"][k]), k
    assert np.array_equal(h["cg_ops"][: int(h["cg_off"][-1])], g["hits"]["cg_ops"][: int(h["cg_off"][-1])])


def test_core_matches_oracle(pafhost, oracle, gold_text):
    for text in (gold_text, gold_text[:-1], paf_cases.ODD, paf_cases.ODD + b"\n", b"", b"\n\n", gold_text[:50000].rsplit(b"\n", 1)[0]):
        ref, n = oracle.parse_paf(text)
        got, m = oracle_ffi.parse_paf_with(pafhost.pafhost_parse, text)
        assert n == m and paf_cases.same_hits(got, ref)
    ref, _ = oracle.parse_paf(paf_cases.ODD)
    assert len(ref["q_id"]) == 6 and ref["mapq"][4] == 300 % 256 and list(ref["is_rev"]) == [0, 1, 0, 0, 0, 1]
    assert list(np.diff(ref["cg_off"])) == [5, 0, 5, 0, 0, 1]


def test_short_line_is_refused(pafhost, oracle):
    bad = b"1\t2\t3\t4\t+\t5\t6\t7\t8\t9\t10\n"
    assert oracle.parse_paf(bad)[1] == -2
    assert oracle_ffi.parse_paf_with(pafhost.pafhost_parse, bad)[1] == -2


line_of = paf_cases.line_of


def test_oracle_number_reading(oracle):
    """The column reader against what `istringstream >> uint32_t` returns in a stand-alone C++ program (libstdc++ 13)."""
    cases = {b" 12": 12, b"+12": 12, b"-5": 4294967291, b"4294967295": 4294967295, b"4294967296": 4294967295, b"99999999999": 4294967295,
             b"12abc": 12, b"abc": 0, b"\r12": 12, b"0012": 12, b"-0": 0, b"- 5": 0, b"+-5": 0, b"1e3": 1, b"0x10": 0,
             b"-4294967295": 1, b"-4294967296": 4294967295}
    for tok, want in cases.items():
        cols = [b"1", tok, b"3", b"4", b"+", b"5", b"6", b"7", b"8", b"9", b"10", b"60"]
        h, n = oracle.parse_paf(line_of(cols) + b"\n")
        assert n == 1 and int(h["q_len"][0]) == want, tok


def test_core_matches_oracle_on_fuzzed_lines(pafhost, oracle):
    """Lines assembled from odd tokens: signs, blanks, overflow, empty columns, carriage returns, tags in any order."""
    text, lines = paf_cases.fuzzed(11)
    ref, n = oracle.parse_paf(text)
    got, m = oracle_ffi.parse_paf_with(pafhost.pafhost_parse, text)
    assert n == m == 4000
    for k in oracle_ffi.PAF_COLS + ("is_rev", "mapq", "cg_off"):
        bad = np.nonzero(got[k] != ref[k])[0]
        assert len(bad) == 0, (k, bad[:3], lines[int(bad[0])] if k != "cg_off" else None)
    assert paf_cases.same_hits(got, ref)
