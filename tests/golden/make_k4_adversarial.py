#!/usr/bin/env python
"""Regenerates tests/golden/k4adv_* by RUNNING THE REFERENCE here (needs /root/reference; never runs on the GPU box).

gen_k4_adversarial.py lays 40 contigs along a line and lets every read cover a window of it, so the same contig pairs are
bridged by 20-40 reads each, with ragged alignment ends (begin / end positions that tie or differ by a few bases between
reads: the `>=` vs `>` optimum rules of the two interval sweeps, ties between begin and end events) and overlapping
consecutive hits. Kept per seed:
  k4adv_<seed>.paf.gz      the PAF
  k4adv_<seed>.npz         contig lengths, mean k-mer counts, read lengths
  k4adv_<seed>.coords.txt  log_coordinate.txt of the reference, reduced (tests/io_helpers.reduce_coordinate_log)
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

REF = os.path.join(ROOT, "oracle", "_ref", "haslr_assemble_ref")
SEEDS, N_READS = (1, 2), 300

for seed in SEEDS:
    tmp = tempfile.mkdtemp(prefix="k4adv_")
    try:
        subprocess.run([sys.executable, os.path.join(HERE, "gen_k4_adversarial.py"), str(seed), tmp, str(N_READS)], check=True)
        with open(os.path.join(tmp, "o.log"), "w") as o, open(os.path.join(tmp, "e.log"), "w") as e:
            subprocess.run([REF, "-t", "1", "-c", "contigs.fa", "-l", "reads.fa", "-m", "map.paf", "-d", "out", "--aln-block", "500",
                            "--aln-sim", "0.85", "--edge-sup", "3"], cwd=tmp, check=True, stdout=o, stderr=e)
        lens, km, kc, _ = io_helpers.load_contigs(os.path.join(tmp, "contigs.fa"))
        reads = io_helpers.load_fasta(os.path.join(tmp, "reads.fa"))
        np.savez_compressed(os.path.join(HERE, f"k4adv_{seed}.npz"), contig_len=lens, mean_kmer=km,
                            read_len=np.array([len(r) for r in reads], dtype=np.uint32))
        with open(os.path.join(tmp, "map.paf"), "rb") as f, gzip.open(os.path.join(HERE, f"k4adv_{seed}.paf.gz"), "wb") as g:
            g.write(f.read())
        with open(os.path.join(HERE, f"k4adv_{seed}.coords.txt"), "w") as f:
            f.write(io_helpers.reduce_coordinate_log(os.path.join(tmp, "out", "log_coordinate.txt")))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
for n in sorted(os.listdir(HERE)):
    if n.startswith("k4adv"):
        print("%9d  %s" % (os.path.getsize(os.path.join(HERE, n)), n))
