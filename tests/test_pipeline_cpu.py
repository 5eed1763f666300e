"""The whole haslr_assemble path on the CPU, against the reference binary's assembly of the adversarial datasets
(tests/golden/k4adv_*): the device stages are played by the oracle (PAF tokeniser, compact reads, edge table, edge
coordinates, POA), everything between and after them is the drop-in binary's own host code (graph container, cleaning, edge
order, segment extraction, stitching). What the GPU drop-in test checks with the kernels in place, without a GPU."""
import ctypes as C
import gzip
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import golden_io
import io_helpers
import oracle_ffi
from test_asm_host import asm, u32p, u64p          # noqa: F401  (fixture)


@pytest.mark.parametrize("seed", [1, 2])
def test_assembly_of_adversarial_dataset_matches_reference(asm, oracle, seed, tmp_path):
    a = golden_io.k4_adversarial(seed)
    with tempfile.TemporaryDirectory() as tmp:      # contigs.fa / reads.fa: the committed generator, same seed (the PAF must come out identical)
        subprocess.run([sys.executable, os.path.join(golden_io.GOLD, "gen_k4_adversarial.py"), str(seed), tmp, str(len(a["read_len"]))],
                       check=True, stdout=subprocess.DEVNULL)
        with open(os.path.join(tmp, "map.paf"), "rb") as f:
            assert f.read() == a["paf"], "generator drifted from the committed fixture"
        contigs = io_helpers.load_fasta(os.path.join(tmp, "contigs.fa"))
        reads = io_helpers.load_fasta(os.path.join(tmp, "reads.fa"))
    hits, _ = oracle.parse_paf(a["paf"])
    n_reads = len(reads)
    read_off = np.searchsorted(hits["q_id"], np.arange(n_reads + 1), side="left").astype(np.uint32)
    elems, off = oracle.compact_lr(hits, read_off, a["mean_kmer"], io_helpers.calc_uniq_freq(a["contig_len"], a["mean_kmer"]))
    key, soff, supp, _ = oracle.backbone_edges(hits["t_id"][elems["hit"]], hits["is_rev"][elems["hit"]], off, 3)
    p = lambda x, t: x.ctypes.data_as(t)
    coff = np.concatenate(([0], np.cumsum([len(c) for c in contigs]))).astype(np.uint64)
    asm.asmhost_prepare.argtypes = [C.c_uint32, C.c_char_p, u64p, C.c_uint64, u64p, u32p, C.c_void_p, C.c_uint32, C.c_char_p]
    n = asm.asmhost_prepare(len(contigs), b"".join(contigs), p(coff, u64p), len(key), p(key, u64p), p(soff, u32p), supp.ctypes.data, 3, str(tmp_path).encode())
    e4 = np.zeros(4 * n, dtype=np.uint32); eso = np.zeros(n + 1, dtype=np.uint32); esupp = np.zeros(len(supp), dtype=oracle_ffi.EDGE_SUPP)
    asm.asmhost_edges.argtypes = [u32p, u32p, C.c_void_p, C.c_uint32]
    ns = asm.asmhost_edges(p(e4, u32p), p(eso, u32p), esupp.ctypes.data, len(esupp))
    e4 = e4.reshape(n, 4); esupp = esupp[:ns]
    oe, os_ = oracle.edge_coords((e4[:, 1] | (e4[:, 3] << 1)).astype(np.uint8), eso, esupp, elems, off, a["read_len"], hits)
    # cns_supp lists (Assemble.cpp:318-326) -> segments (the binary's own extraction) -> consensus (oracle POA)
    cns, edge_seg_off = [], [0]
    for e in range(n):
        for i in range(int(eso[e]), int(eso[e + 1])):
            o = os_[i]
            if o["in_best"] and o["lr_start"] != -1 and o["lr_end"] != -1:
                cns.append((int(esupp[i]["lr_id_strand"]) & 0x7FFFFFFF, int(o["lr_strand"]), int(o["lr_start"]) + 1, int(o["lr_end"]) - 1))
        edge_seg_off.append(len(cns))
    cns4 = np.array(cns, dtype=np.uint32).reshape(-1)
    roff = np.concatenate(([0], np.cumsum([len(r) for r in reads]))).astype(np.uint64)
    cap = int(sum(c[3] - c[2] + 1 for c in cns)) + 1024
    seg = C.create_string_buffer(cap); seg_off = np.zeros(len(cns) + 1, dtype=np.uint64)
    asm.asmhost_segments.restype = C.c_longlong
    asm.asmhost_segments.argtypes = [C.c_char_p, u64p, C.c_uint32, u32p, C.c_uint32, C.c_char_p, C.c_uint64, u64p]
    nb = asm.asmhost_segments(b"".join(reads), p(roff, u64p), n_reads, p(cns4, u32p), len(cns), seg, cap, p(seg_off, u64p))
    assert nb > 0
    cons, coffs, _, _ = oracle.poa_batch(np.frombuffer(seg.raw[:nb], dtype=np.uint8), seg_off, np.array(edge_seg_off, dtype=np.uint32), threads=8)
    lens = a["contig_len"].astype(np.int64)
    he = np.where(oe["n_cns"] > 0, oe["c1"], np.where(e4[:, 1] == 0, lens[e4[:, 0]] - 1, 0)).astype(np.uint32)
    tb = np.where(oe["n_cns"] > 0, oe["c2"], np.where(e4[:, 3] == 0, 0, lens[e4[:, 2]] - 1)).astype(np.uint32)
    out = tmp_path / "out"; out.mkdir()
    asm.asmhost_finish.argtypes = [u32p, u32p, u32p, C.c_char_p, u64p, C.c_char_p]
    ncns = np.ascontiguousarray(oe["n_cns"], dtype=np.uint32)
    coffs = np.ascontiguousarray(coffs, dtype=np.uint64)
    assert asm.asmhost_finish(p(he, u32p), p(tb, u32p), p(ncns, u32p), cons.tobytes(), p(coffs, u64p), str(out).encode()) == 0
    with open(out / "asm.final.fa", "rb") as f, gzip.open(os.path.join(golden_io.GOLD, f"k4adv_{seed}.asm.final.fa.gz"), "rb") as gz:
        assert f.read() == gz.read()
    with open(out / "asm.final.ann") as f:
        assert f.read() == golden_io.text(f"k4adv_{seed}.asm.final.ann")
