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


def coords():
    """[dict(edge=(node1, rev1, node2, rev2), n_supp, int1, int2, c1, c2, n_best, detail=[(hs, he, hrev, ts, te, trev)],
    best=[(lr, len, strand, spos, epos) | (lr, len, strand, None, None)])] in the order the reference processed the edges."""
    out = []
    with open(os.path.join(GOLD, "syn200k_coords.txt")) as f:
        for line in f:
            t = line.split()
            if t[0] == "E":
                v = list(map(int, t[1:]))
                out.append(dict(edge=tuple(v[0:4]), n_supp=v[4], int1=(v[5], v[6]), int2=(v[7], v[8]), c1=v[9], c2=v[10], n_best=v[11],
                                detail=[], best=[]))
            elif t[0] == "D":
                out[-1]["detail"].append(tuple(map(int, t[1:])))
            elif t[0] == "S":
                out[-1]["best"].append((int(t[1]), int(t[2]), int(t[3])) + ((None, None) if t[4] == "X" else (int(t[4]), int(t[5]))))
    return out


def read_len():
    return np.load(os.path.join(GOLD, "syn200k_read_len.npy"))


def coord_inputs(oracle):
    """Inputs of the edge-coordinate stage for the golden dataset: the edges the reference processed (after cleaning), with
    their supports looked up in the oracle's K2 table of the oracle's K1 output (cleaning removes edges, never edits them)."""
    g = inputs()
    elems, off = oracle.compact_lr(g["hits"], g["read_off"], g["mean_kmer"], g["uniq_freq"])
    key, soff, supp, _ = oracle.backbone_edges(g["hits"]["t_id"][elems["hit"]], g["hits"]["is_rev"][elems["hit"]], off, 3)
    pos = {int(k): i for i, k in enumerate(key.tolist())}
    gold = coords()
    rev, so, parts = [], [0], []
    for e in gold:
        n1, r1, n2, r2 = e["edge"]
        i = pos[(((n1 << 1) | r1) << 32) | ((n2 << 1) | r2)]
        parts.append(supp[int(soff[i]): int(soff[i + 1])])
        so.append(so[-1] + len(parts[-1]))
        rev.append(r1 | (r2 << 1))
    return dict(g=g, elems=elems, cl_off=off, edge_rev=np.array(rev, dtype=np.uint8), supp_off=np.array(so, dtype=np.uint32),
                supp=np.concatenate(parts), read_len=read_len(), gold=gold)


def k1_adversarial(seed):
    """Adversarial K1/K2 fixture (tests/golden/make_k1_adversarial.py): PAF text, contig table, reference compact_uniq.txt / GFA links."""
    z = np.load(os.path.join(GOLD, f"k1adv_{seed}.npz"))
    with gzip.open(os.path.join(GOLD, f"k1adv_{seed}.paf.gz"), "rb") as f:
        paf = f.read()
    return dict(paf=paf, contig_len=z["contig_len"], mean_kmer=np.ascontiguousarray(z["mean_kmer"]), n_reads=int(z["n_reads"]),
                compact=text(f"k1adv_{seed}.compact_uniq.txt"), links01=text(f"k1adv_{seed}.links01"), links02=text(f"k1adv_{seed}.links02"))


def parse_coords_text(txt):
    out = []
    for line in txt.splitlines():
        t = line.split()
        if t[0] == "E":
            v = list(map(int, t[1:]))
            out.append(dict(edge=tuple(v[0:4]), n_supp=v[4], int1=(v[5], v[6]), int2=(v[7], v[8]), c1=v[9], c2=v[10], n_best=v[11], detail=[], best=[]))
        elif t[0] == "D":
            out[-1]["detail"].append(tuple(map(int, t[1:])))
        elif t[0] == "S":
            out[-1]["best"].append((int(t[1]), int(t[2]), int(t[3])) + ((None, None) if t[4] == "X" else (int(t[4]), int(t[5]))))
    return out


def k4_adversarial(seed):
    """Adversarial K4 fixture (tests/golden/make_k4_adversarial.py): PAF text, contig / read tables, the reference's reduced log_coordinate.txt."""
    z = np.load(os.path.join(GOLD, f"k4adv_{seed}.npz"))
    with gzip.open(os.path.join(GOLD, f"k4adv_{seed}.paf.gz"), "rb") as f:
        paf = f.read()
    return dict(paf=paf, contig_len=z["contig_len"], mean_kmer=np.ascontiguousarray(z["mean_kmer"]), read_len=z["read_len"],
                gold=parse_coords_text(text(f"k4adv_{seed}.coords.txt")))
