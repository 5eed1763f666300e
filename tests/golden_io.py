"""Loaders for tests/golden/* (fixtures made by tests/golden/make_golden.py from the reference binary)."""
import gzip
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def inputs():
    z = np.load(os.path.join(GOLD, "syn200k_inputs.npz"))
    hits = {k[4:]: np.ascontiguousarray(z[k]) for k in z.files if k.startswith("hit_")}
    return dict(hits=hits, read_off=np.ascontiguousarray(z["read_off"]), mean_kmer=np.ascontiguousarray(z["mean_kmer"]),
                contig_len=z["contig_len"], uniq_freq=float(z["uniq_freq"]), n_reads=int(z["n_reads"]))


def text(name):
    with open(os.path.join(GOLD, name)) as f:
        return f.read()


def poa_edges():
    """[(label, [segment bytes...], consensus bytes)] in the order the reference processed them."""
    out = []
    with gzip.open(os.path.join(GOLD, "syn200k_poa.txt.gz"), "rt") as f:
        for line in f:
            tag, _, rest = line.partition(" ")
            rest = rest.rstrip("\n")
            if tag == "E":
                out.append([rest, [], b""])
            elif tag == "S":
                out[-1][1].append(rest.encode())
            elif tag == "C":
                out[-1][2] = rest.encode()
    return out
