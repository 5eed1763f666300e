#!/usr/bin/env python
"""Regenerates tests/golden/* by RUNNING THE REFERENCE here (needs /root/reference; never runs on the GPU box).

  1. oracle/_ref/gen_synth writes a seeded synthetic dataset (contigs.fa, reads.fa, map.paf)
  2. oracle/_ref/haslr_assemble_ref — the unmodified reference sources compiled by oracle/Makefile, with
     oracle/spoa_restated/spoa.hpp standing in for the un-vendored SPOA — runs on it with haslr.py's flags
  3. harvested into small fixtures:
       syn200k_inputs.npz          parsed PAF hits (SoA), read offsets, contig lengths / mean k-mer, uniq_freq
       syn200k_compact_uniq.txt    reference output #1 (print_compact_longreads, Longread.cpp:675-693)
       syn200k_backbone01.links    L lines of backbone.01.init.gfa   (bbg_print_graph_gfa, Backbone_graph.cpp:540-588)
       syn200k_backbone02.links    L lines of backbone.02.weakEdge.gfa
       syn200k_backbone0[1-6].links/.stat  the same after every cleaning stage (tips, simple / super / small bubbles: Cleaning.cpp)
       syn200k_backbone.0*.log, .branching.log  the cleaning logs
       syn200k_poa.txt.gz          log_consensus.txt reduced to: per edge the segments fed to SPOA and the consensus
                                   (the reference's own call sequence, Assemble.cpp:499-554, around the restated SPOA)
       syn200k_coords.txt          log_coordinate.txt reduced to: per edge  E node1 rev1 node2 rev2 n_supp int1 int2 c1 c2 n_best,
                                   per support  D head(t_start t_end strand) tail(...),  per best support  S lr len strand
                                   spos epos | X   (asm_calc_single_edge_coordinates, Assemble.cpp:157-363)
       syn200k_read_len.npy        long-read lengths (Longread_List_t::reads[i].len)
       syn200k_asm.final.fa.gz, .ann  the assembly and its annotation (stitching, Assemble.cpp:607-810,1045-1077)
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import io_helpers  # noqa: E402

GEN = os.path.join(ROOT, "oracle", "_ref", "gen_synth")
REF = os.path.join(ROOT, "oracle", "_ref", "haslr_assemble_ref")
GENOME, N_READS, READ_LEN, SEED = 200000, 600, 8000, 7


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "tools", "ref"], check=True)
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        subprocess.run([GEN, tmp, str(GENOME), str(N_READS), str(READ_LEN), str(SEED)], check=True, stdout=subprocess.DEVNULL)
        with open(os.path.join(tmp, "out.log"), "w") as o, open(os.path.join(tmp, "err.log"), "w") as e:
            subprocess.run([REF, "-t", "1", "-c", "contigs.fa", "-l", "reads.fa", "-m", "map.paf", "-d", "out",
                            "--aln-block", "500", "--aln-sim", "0.85", "--edge-sup", "3"], cwd=tmp, check=True, stdout=o, stderr=e)
        out = os.path.join(tmp, "out")
        lens, km, kc, _ = io_helpers.load_contigs(os.path.join(tmp, "contigs.fa"))
        hits, read_off = io_helpers.parse_paf(os.path.join(tmp, "map.paf"), N_READS)
        np.savez_compressed(os.path.join(HERE, "syn200k_inputs.npz"), contig_len=lens, mean_kmer=km,
                            uniq_freq=np.float64(io_helpers.calc_uniq_freq(lens, km)), n_reads=np.uint32(N_READS),
                            read_off=read_off, **{"hit_" + k: v for k, v in hits.items()})
        shutil.copy(os.path.join(out, "compact_uniq.txt"), os.path.join(HERE, "syn200k_compact_uniq.txt"))
        for tag, name in (("01", "backbone.01.init"), ("02", "backbone.02.weakEdge"), ("03", "backbone.03.tip"),
                          ("04", "backbone.04.simplebubble"), ("05", "backbone.05.superbubble"), ("06", "backbone.06.smallbubble")):
            with open(os.path.join(HERE, f"syn200k_backbone{tag}.links"), "w") as f:
                f.write(io_helpers.gfa_links_of_file(os.path.join(out, name + ".gfa")))
            shutil.copy(os.path.join(out, name + ".stat"), os.path.join(HERE, f"syn200k_backbone{tag}.stat"))
        for name in ("backbone.03.tip.log", "backbone.04.simplebubble.log", "backbone.05.superbubble.log", "backbone.06.smallbubble.log",
                     "backbone.branching.log"):
            shutil.copy(os.path.join(out, name), os.path.join(HERE, "syn200k_" + name))
        # POA: segments and consensus per edge, in the order the reference fed them
        with open(os.path.join(out, "log_consensus.txt")) as f, gzip.open(os.path.join(HERE, "syn200k_poa.txt.gz"), "wt") as g:
            want_seq = False
            for line in f:
                if line.startswith("calc_cns"):
                    g.write("E " + line.split(" ", 2)[2])
                elif line.startswith(">CONSENSUS"):
                    want_seq = "C"
                elif line.startswith(">"):
                    want_seq = "S"
                elif want_seq:
                    g.write(want_seq + " " + line)
                    want_seq = False
        # edge coordinates
        with open(os.path.join(HERE, "syn200k_coords.txt"), "w") as g:
            g.write(io_helpers.reduce_coordinate_log(os.path.join(out, "log_coordinate.txt")))
        # the assembly itself (stitching: Assemble.cpp:607-810,1045-1077)
        with open(os.path.join(out, "asm.final.fa"), "rb") as f, gzip.open(os.path.join(HERE, "syn200k_asm.final.fa.gz"), "wb") as g:
            g.write(f.read())
        shutil.copy(os.path.join(out, "asm.final.ann"), os.path.join(HERE, "syn200k_asm.final.ann"))
        reads = io_helpers.load_fasta(os.path.join(tmp, "reads.fa"))
        np.save(os.path.join(HERE, "syn200k_read_len.npy"), np.array([len(r) for r in reads], dtype=np.uint32))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for n in sorted(os.listdir(HERE)):
        print("%9d  %s" % (os.path.getsize(os.path.join(HERE, n)), n))


CLEAN_GENOME, CLEAN_READS, CLEAN_SEED = 2000000, 6000, 5
CLEAN_FILES = ["backbone.%s.%s" % (a, e) for a in ("01.init", "02.weakEdge", "03.tip", "04.simplebubble", "05.superbubble", "06.smallbubble")
               for e in ("stat", "gfa")] + ["backbone.03.tip.log", "backbone.04.simplebubble.log", "backbone.05.superbubble.log",
                                            "backbone.06.smallbubble.log", "backbone.branching.log"]


def make_cleaning_golden():
    """syn2m_cleaning.json.gz: every backbone.* file of the reference on a 2 Mb dataset (GFA reduced to its L lines) — tips,
    simple / super / small bubbles all occur there; the dataset itself is regenerated by the tests (tools/gen_synth.cpp, seed 5)."""
    import json
    tmp = tempfile.mkdtemp(prefix="golden_clean_")
    try:
        subprocess.run([GEN, tmp, str(CLEAN_GENOME), str(CLEAN_READS), str(READ_LEN), str(CLEAN_SEED)], check=True, stdout=subprocess.DEVNULL)
        with open(os.path.join(tmp, "out.log"), "w") as o, open(os.path.join(tmp, "err.log"), "w") as e:
            subprocess.run([REF, "-t", "1", "-c", "contigs.fa", "-l", "reads.fa", "-m", "map.paf", "-d", "out",
                            "--aln-block", "500", "--aln-sim", "0.85", "--edge-sup", "3"], cwd=tmp, check=True, stdout=o, stderr=e)
        bundle = {}
        for name in CLEAN_FILES:
            path = os.path.join(tmp, "out", name)
            if name.endswith(".gfa"):
                bundle[name + ".links"] = io_helpers.gfa_links_of_file(path)
            else:
                with open(path) as f:
                    bundle[name] = f.read()
        with gzip.open(os.path.join(HERE, "syn2m_cleaning.json.gz"), "wt") as g:
            json.dump(bundle, g)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
    make_cleaning_golden()
    for n in sorted(os.listdir(HERE)):
        print("%9d  %s" % (os.path.getsize(os.path.join(HERE, n)), n))
