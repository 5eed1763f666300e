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


# ---------------------------------------------------------------------------------------------------------------------
# cfg3-shaped batches from a counter-based generator: the same bytes whether they are made with numpy on the host or with
# torch on a GPU, edge by edge (edge e depends on (seed, e) only), so bench.py's GPU arm and its CPU reference arm can be
# fed identical segments and any prefix of a shard is a valid sample of it.
# ---------------------------------------------------------------------------------------------------------------------
def _const64(x, np_backend):
    return np.uint64(x) if np_backend else (x - (1 << 64) if x >= (1 << 63) else x)


def _mix64(z, np_backend):
    """splitmix64 finaliser on uint64 (numpy) / int64 with wrap-around (torch): identical bit patterns."""
    def shr(v, k):
        return v >> np.uint64(k) if np_backend else (v >> k) & ((1 << (64 - k)) - 1)
    z = z + _const64(0x9E3779B97F4A7C15, np_backend)
    z = (z ^ shr(z, 30)) * _const64(0xBF58476D1CE4E5B9, np_backend)
    z = (z ^ shr(z, 27)) * _const64(0x94D049BB133111EB, np_backend)
    return z ^ shr(z, 31)


def hashed_batch(n_edges, seed, depth=6, length=1500, err=(0.04, 0.03, 0.02), device=None, edge0=0, chunk=4096, eids=None):
    """Per edge: truth of `length` bases, `depth` copies with per-base deletion / substitution and single-base insertions (the
    error model of poa_batch / SURVEY 8(d) cfg3). device None: numpy; else a torch device (the bases stay there).
    eids: explicit edge ids (any subset, any order) instead of edge0 .. edge0 + n_edges.
    Returns (bases, seg_off uint64 numpy, edge_seg_off uint32 numpy)."""
    npb = device is None
    if eids is not None:
        eids = np.asarray(eids, dtype=np.int64)
        n_edges = len(eids)
    if not npb:
        import torch
    p_ins, p_del, p_sub = err
    t_del, t_sub, t_ins = int(p_del * 2**32), int((p_del + p_sub) * 2**32), int(p_ins * 2**32)
    parts, lens_all = [], []
    old = np.seterr(over="ignore") if npb else None
    try:
        for a in range(0, n_edges, chunk):
            e = min(chunk, n_edges - a)
            if npb:
                eid = (np.arange(e, dtype=np.uint64) + np.uint64(edge0 + a) if eids is None else eids[a: a + e].astype(np.uint64))[:, None, None]
                rd = np.arange(depth, dtype=np.uint64)[None, :, None]
                pos = np.arange(length, dtype=np.uint64)[None, None, :]
                m32 = np.uint64(0xFFFFFFFF)
                sh = lambda v, k: v << np.uint64(k)
                u32 = lambda v: (v >> np.uint64(32)) & m32
            else:
                eid = (torch.arange(e, dtype=torch.int64, device=device) + (edge0 + a) if eids is None
                       else torch.from_numpy(eids[a: a + e]).to(device))[:, None, None]
                rd = torch.arange(depth, dtype=torch.int64, device=device)[None, :, None]
                pos = torch.arange(length, dtype=torch.int64, device=device)[None, None, :]
                m32 = 0xFFFFFFFF
                sh = lambda v, k: v << k
                u32 = lambda v: (v >> 32) & m32
            ekey = _mix64(sh(eid, 20) + _const64((seed * 0x2545F4914F6CDD1D) & ((1 << 64) - 1), npb), npb)      # [e,1,1]
            truth = _mix64(ekey + pos, npb) & _const64(3, npb)                             # [e,1,L]
            rkey = _mix64(ekey ^ sh(rd + _const64(1, npb), 40), npb)                        # [e,D,1]
            h1 = _mix64(rkey + sh(pos, 2), npb)
            h2 = _mix64(rkey + sh(pos, 2) + _const64(1, npb), npb)
            u = u32(h1)                                                                     # keep / substitute / delete
            keep = u >= t_del
            sub = keep & (u < t_sub)
            shift = (h1 & _const64(0xFFFF, npb)) % _const64(3, npb) + _const64(1, npb)
            code = (truth + shift * sub) & _const64(3, npb)
            ins = u32(h2) < t_ins
            ins_code = h2 & _const64(3, npb)
            if npb:
                keep2 = keep.reshape(e * depth, length); ins2 = ins.reshape(e * depth, length)
                code2 = np.broadcast_to(code, (e, depth, length)).reshape(e * depth, length); ic2 = ins_code.reshape(e * depth, length)
                cnt = keep2.astype(np.int64) + ins2.astype(np.int64)
                lens = cnt.sum(1)
                end = np.cumsum(cnt, 1) + (np.cumsum(lens) - lens)[:, None]
                out = np.empty(int(lens.sum()), dtype=np.uint8)
                out[(end - cnt)[keep2]] = ACGT[code2[keep2].astype(np.int64)]
                out[(end - 1)[ins2]] = ACGT[ic2[ins2].astype(np.int64)]
                parts.append(out); lens_all.append(lens)
            else:
                acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
                keep2 = keep.reshape(e * depth, length); ins2 = ins.reshape(e * depth, length)
                code2 = code.expand(e, depth, length).reshape(e * depth, length); ic2 = ins_code.reshape(e * depth, length)
                cnt = keep2.to(torch.int64) + ins2.to(torch.int64)
                lens = cnt.sum(1)
                end = torch.cumsum(cnt, 1) + (torch.cumsum(lens, 0) - lens)[:, None]
                out = torch.empty(int(lens.sum().item()), dtype=torch.uint8, device=device)
                out[(end - cnt)[keep2]] = acgt[code2[keep2]]
                out[(end - 1)[ins2]] = acgt[ic2[ins2]]
                parts.append(out); lens_all.append(lens.cpu().numpy())
    finally:
        if npb:
            np.seterr(**old)
    if npb:
        bases = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    else:
        bases = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.uint8, device=device)
    lens = np.concatenate(lens_all).astype(np.uint64) if lens_all else np.zeros(0, np.uint64)
    seg_off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
    edge_seg_off = (np.arange(n_edges + 1, dtype=np.uint64) * depth).astype(np.uint32)
    return bases, seg_off, edge_seg_off
