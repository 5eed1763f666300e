#!/usr/bin/env python
"""Regenerates tests/golden/k1adv_* by RUNNING THE REFERENCE here (needs /root/reference; never runs on the GPU box).

gen_k1_adversarial.py writes small datasets whose PAF stresses what the regular generator barely touches: consecutive hits
overlapping on the read by anything from 1 bp to whole hits, identical starts, identical ends (ties of the per-read sort on
(q_end, q_start)), repeated hits of one contig, indels next to the trimmed ends, rows at the filter thresholds. The reference
binary (oracle/_ref/haslr_assemble_ref) is run on each; kept per seed:
  k1adv_<seed>.paf.gz            the PAF
  k1adv_<seed>.npz               contig lengths and mean k-mer counts
  k1adv_<seed>.compact_uniq.txt  reference output (print_compact_longreads, Longread.cpp:675-693)
  k1adv_<seed>.links0[12]        L lines of backbone.01.init.gfa / backbone.02.weakEdge.gfa
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
SEEDS, N_READS = (1, 2), 200

for seed in SEEDS:
    tmp = tempfile.mkdtemp(prefix="k1adv_")
    try:
        subprocess.run([sys.executable, os.path.join(HERE, "gen_k1_adversarial.py"), str(seed), tmp, str(N_READS)], check=True)
        with open(os.path.join(tmp, "o.log"), "w") as o, open(os.path.join(tmp, "e.log"), "w") as e:
            subprocess.run([REF, "-t", "1", "-c", "contigs.fa", "-l", "reads.fa", "-m", "map.paf", "-d", "out", "--aln-block", "500",
                            "--aln-sim", "0.85", "--edge-sup", "3"], cwd=tmp, check=True, stdout=o, stderr=e)
        lens, km, kc, _ = io_helpers.load_contigs(os.path.join(tmp, "contigs.fa"))
        np.savez_compressed(os.path.join(HERE, f"k1adv_{seed}.npz"), contig_len=lens, mean_kmer=km, n_reads=np.uint32(N_READS))
        with open(os.path.join(tmp, "map.paf"), "rb") as f, gzip.open(os.path.join(HERE, f"k1adv_{seed}.paf.gz"), "wb") as g:
            g.write(f.read())
        shutil.copy(os.path.join(tmp, "out", "compact_uniq.txt"), os.path.join(HERE, f"k1adv_{seed}.compact_uniq.txt"))
        for tag, name in (("01", "backbone.01.init"), ("02", "backbone.02.weakEdge")):
            with open(os.path.join(HERE, f"k1adv_{seed}.links{tag}"), "w") as f:
                f.write(io_helpers.gfa_links_of_file(os.path.join(tmp, "out", name + ".gfa")))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
for n in sorted(os.listdir(HERE)):
    if n.startswith("k1adv"):
        print("%9d  %s" % (os.path.getsize(os.path.join(HERE, n)), n))
