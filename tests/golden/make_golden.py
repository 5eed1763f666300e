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
       syn200k_backbone0[12].stat  bbg_general_stats output
       syn200k_poa.txt.gz          log_consensus.txt reduced to: per edge the segments fed to SPOA and the consensus
                                   (the reference's own call sequence, Assemble.cpp:499-554, around the restated SPOA)
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
        for tag, name in (("01", "backbone.01.init"), ("02", "backbone.02.weakEdge")):
            with open(os.path.join(HERE, f"syn200k_backbone{tag}.links"), "w") as f:
                f.write(io_helpers.gfa_links_of_file(os.path.join(out, name + ".gfa")))
            shutil.copy(os.path.join(out, name + ".stat"), os.path.join(HERE, f"syn200k_backbone{tag}.stat"))
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
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for n in sorted(os.listdir(HERE)):
        print("%9d  %s" % (os.path.getsize(os.path.join(HERE, n)), n))


if __name__ == "__main__":
    main()
