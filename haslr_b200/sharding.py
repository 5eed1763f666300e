"""Multi-GPU plumbing of the path: backbone edges are independent, so they are dealt to ranks by estimated DP cost and
every rank runs hgpu_poa_batch on its own shard; the only exchange is one variable-length all-gather of the per-shard
consensus at the end. torch.distributed is plumbing here: backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
import numpy as np


STRIPE = 512          # columns of one score-matrix stripe (csrc/poa_device.cuh: DP_NW16 x 2 x 32)
GRAPH_ROWS = 3.0      # serial graph work per node, in units of one stripe row (csrc/poa.cu: POOL_GRAPH_ROWS)


def alignment_cost(v, l):
    """Estimated time of aligning a segment of l bases to a graph of v nodes, in stripe-row units: the fill computes whole
    512-column stripes whatever the gap is, and traceback / graph update / topological sort cost ~3 stripe rows per node."""
    return (v + 1.0) * (np.ceil((l + 1.0) / STRIPE) + GRAPH_ROWS)


def edge_costs(seg_off, edge_seg_off):
    """Estimated time per edge from the segment lengths alone: sum over alignments of alignment_cost, |V| growing ~10 % per read
    (the model csrc/poa.cu uses to share the device between size classes)."""
    seg_off = np.asarray(seg_off, dtype=np.int64)
    eso = np.asarray(edge_seg_off, dtype=np.int64)
    lens = np.diff(seg_off)
    cost = np.zeros(len(eso) - 1, dtype=np.float64)
    for e in range(len(eso) - 1):
        v = 0.0
        c = 0.0
        for k, l in enumerate(lens[eso[e]: eso[e + 1]]):
            if k:
                c += alignment_cost(v, float(l))
            v = max(v, float(l)) + 0.1 * l
        cost[e] = c
    return cost


def shard_edges(cost, world):
    """Longest-processing-time-first deal of edges to `world` ranks. Returns a list of sorted index arrays (a partition)."""
    order = np.lexsort((np.arange(len(cost)), -np.asarray(cost, dtype=np.float64)))
    load = np.zeros(world, dtype=np.float64)
    out = [[] for _ in range(world)]
    for e in order:
        r = int(np.argmin(load))
        out[r].append(int(e))
        load[r] += cost[e]
    return [np.array(sorted(x), dtype=np.int64) for x in out]


def take_shard(bases, seg_off, edge_seg_off, edges):
    """Sub-batch (bases, seg_off, edge_seg_off) holding only `edges`, in that order."""
    seg_off = np.asarray(seg_off, dtype=np.uint64)
    eso = np.asarray(edge_seg_off, dtype=np.uint32)
    segs = np.concatenate([np.arange(eso[e], eso[e + 1]) for e in edges]).astype(np.int64) if len(edges) else np.zeros(0, np.int64)
    lens = (seg_off[segs + 1] - seg_off[segs]).astype(np.uint64) if len(segs) else np.zeros(0, np.uint64)
    new_off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
    parts = [np.asarray(bases[int(seg_off[s]): int(seg_off[s + 1])]) for s in segs]
    new_bases = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    cnt = np.array([int(eso[e + 1] - eso[e]) for e in edges], dtype=np.int64)
    new_eso = np.concatenate(([0], np.cumsum(cnt))).astype(np.uint32)
    return new_bases.astype(np.uint8), new_off, new_eso


def all_gather_consensus(dist, cons, off, device, to_host=True):
    """One variable-length all-gather: every rank ends with every rank's (consensus bytes, offsets).
    cons: uint8 torch tensor on `device`; off: numpy uint64[n_local+1]. Returns list over ranks of (bytes, off): numpy arrays,
    or (to_host=False) the gathered device tensors, trimmed."""
    import torch
    world = dist.get_world_size()
    n_local = len(off) - 1
    meta = torch.tensor([int(off[-1]), n_local], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_b = max(int(m[0]) for m in metas)
    max_n = max(int(m[1]) for m in metas)
    pad = torch.zeros(max(max_b, 1), dtype=torch.uint8, device=device)
    pad[: int(off[-1])] = cons[: int(off[-1])]
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    offp = torch.zeros(max_n + 1, dtype=torch.int64, device=device)
    offp[: n_local + 1] = torch.from_numpy(np.asarray(off, dtype=np.int64)).to(device)
    offs = [torch.empty_like(offp) for _ in range(world)]
    dist.all_gather(offs, offp)
    out = []
    for r in range(world):
        nb, ne = int(metas[r][0]), int(metas[r][1])
        if to_host:
            out.append((bufs[r][:nb].cpu().numpy(), offs[r][: ne + 1].cpu().numpy().astype(np.uint64)))
        else:
            out.append((bufs[r][:nb], offs[r][: ne + 1]))
    return out
