"""Seeded synthetic POA inputs (SURVEY.md §8(d) cfg3 shape): per edge a random truth and R noisy copies."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def mutate(rng, truth, p_ins=0.04, p_del=0.03, p_sub=0.02):
    """PacBio-like errors. Vectorised: per truth base one of keep/sub/del, plus geometric-free single insertions."""
    n = len(truth)
    u = rng.random(n)
    keep = u >= p_del
    sub = (u >= p_del) & (u < p_del + p_sub)
    out = truth.copy()
    if sub.any():
        out[sub] = ACGT[(np.searchsorted(ACGT, out[sub]) + rng.integers(1, 4, sub.sum())) % 4]
    ins = rng.random(n) < p_ins
    # interleave: each kept base optionally followed by one inserted base
    cnt = keep.astype(np.int64) + ins.astype(np.int64)
    off = np.concatenate(([0], np.cumsum(cnt)))
    res = np.empty(off[-1], dtype=np.uint8)
    res[off[:-1][keep]] = out[keep]
    res[(off[1:] - 1)[ins]] = ACGT[rng.integers(0, 4, ins.sum())]
    return res


def poa_batch(seed, n_edges, depth=6, length=1500, err=(0.04, 0.03, 0.02), length_jitter=0.0, depth_jitter=0):
    """Returns (bases uint8[], seg_off uint64[], edge_seg_off uint32[], truths list)."""
    rng = np.random.default_rng(seed)
    segs, edge_seg_off, truths = [], [0], []
    for _ in range(n_edges):
        L = length if length_jitter == 0 else max(1, int(rng.normal(length, length * length_jitter)))
        truth = ACGT[rng.integers(0, 4, L)]
        truths.append(truth)
        d = depth if depth_jitter == 0 else max(0, depth + int(rng.integers(-depth_jitter, depth_jitter + 1)))
        for _ in range(d):
            segs.append(mutate(rng, truth, *err))
        edge_seg_off.append(len(segs))
    lens = np.array([len(s) for s in segs], dtype=np.uint64)
    seg_off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
    bases = np.concatenate(segs) if segs else np.zeros(0, dtype=np.uint8)
    return bases, seg_off, np.array(edge_seg_off, dtype=np.uint32), truths


def from_strings(edges):
    """edges: list of lists of bytes -> (bases, seg_off, edge_seg_off)."""
    segs, eo = [], [0]
    for e in edges:
        for s in e:
            segs.append(np.frombuffer(s, dtype=np.uint8))
        eo.append(len(segs))
    lens = np.array([len(s) for s in segs], dtype=np.uint64)
    seg_off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
    bases = np.concatenate(segs) if segs and seg_off[-1] > 0 else np.zeros(0, dtype=np.uint8)
    return bases, seg_off, np.array(eo, dtype=np.uint32)
