"""CPU check of the drop-in binary's FASTA/FASTQ loader (haslr_b200/host/io.cpp): multi-line records, FASTQ, gzip, the
2-bit folding of the reference (Compressed_sequence.cpp:10-19,57: acgt -> ACGT, everything else -> A), contig header tags
(Contig.cpp:63-66) and calc_uniq_freq (Contig.cpp:162-174), against plain Python."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

import io_helpers

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "io_host_check.cpp")
LIB = os.path.join(HERE, "native", "libiotest.so")
PROD = os.path.join(ROOT, "haslr_b200")


@pytest.fixture(scope="module")
def io():
    if not os.path.exists(os.path.join(PROD, "libhaslr_b200.so")):
        pytest.skip("product library not built")
    deps = [SRC, os.path.join(PROD, "host", "io.cpp"), os.path.join(PROD, "host", "haslr.hpp")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", LIB, SRC, os.path.join(PROD, "host", "io.cpp"),
                        "-L" + PROD, "-lhaslr_b200", "-Wl,-rpath," + PROD, "-lz"], check=True)
    L = C.CDLL(LIB)
    L.iohost_load_fasta.restype = C.c_longlong
    L.iohost_load_fasta.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.c_ulonglong,
                                    C.POINTER(C.c_uint), C.POINTER(C.c_double)]
    L.iohost_uniq_freq.restype = C.c_double
    L.iohost_uniq_freq.argtypes = [C.c_char_p]
    return L


def load(L, path, meta=False, cap=1 << 22):
    seq = C.create_string_buffer(cap)
    off = (C.c_ulonglong * 4096)()
    kc = (C.c_uint * 4096)(); km = (C.c_double * 4096)()
    n = L.iohost_load_fasta(path.encode(), int(meta), seq, cap, off, 4096, kc, km)
    assert n >= 0
    recs = [seq.raw[off[i]: off[i + 1]] for i in range(n)]
    return recs, list(kc[:n]), list(km[:n])


def fold(s):
    return bytes(c if c in b"ACGT" else 65 for c in s.upper().replace(b" ", b""))


def test_fasta_fastq_gzip_and_folding(io, tmp_path):
    rng = np.random.default_rng(1)
    raw = [rng.choice(np.frombuffer(b"ACGTacgtNnRYx-", np.uint8), int(rng.integers(1, 400))).tobytes() for _ in range(40)] + [b"A", b"acgtn" * 50]
    fa = tmp_path / "a.fa"
    with open(fa, "wb") as f:
        for i, s in enumerate(raw):
            f.write(b">%d some comment\n" % i)
            for k in range(0, len(s), 60):
                f.write(s[k: k + 60] + (b"\r\n" if i % 7 == 0 else b"\n"))
            if i % 5 == 0:
                f.write(b"\n")
    want = [fold(s) for s in raw]
    assert load(io, str(fa))[0] == want
    gz = tmp_path / "a.fa.gz"
    with open(fa, "rb") as f, gzip.open(gz, "wb") as g:
        g.write(f.read())
    assert load(io, str(gz))[0] == want
    fq = tmp_path / "a.fq"
    with open(fq, "wb") as f:
        for i, s in enumerate(raw):
            f.write(b"@%d\n%s\n+\n%s\n" % (i, s, b"@" * len(s)))        # quality lines that start with '@' must not start a record
    assert load(io, str(fq))[0] == want


def test_contig_tags_and_uniq_freq(io, tmp_path):
    rng = np.random.default_rng(2)
    fa = tmp_path / "c.fa"
    lens, kms = [], []
    with open(fa, "wb") as f:
        for i in range(57):
            n = int(rng.integers(50, 3000)); km = float(np.round(rng.normal(30, 5), 3))
            lens.append(n); kms.append(km)
            f.write(b">%d LN:i:%d KC:i:%d km:f:%.3f\n%s\n" % (i, n, int(km * n), km, rng.choice(np.frombuffer(b"ACGT", np.uint8), n).tobytes()))
    recs, kc, km = load(io, str(fa), meta=True)
    assert [len(r) for r in recs] == lens and kc == [int(k * n) for k, n in zip(kms, lens)] and np.allclose(km, kms)
    assert io.iohost_uniq_freq(str(fa).encode()) == pytest.approx(io_helpers.calc_uniq_freq(np.array(lens, dtype=np.uint32), np.array(kms)), rel=0, abs=1e-12)
