"""Host-side multi-GPU logic on CPU: cost-balanced edge sharding and the variable-length all-gather of consensus,
run with world_size 2 over gloo (the NCCL path in bench.py uses the same functions)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from haslr_b200 import sharding


def test_shard_is_a_balanced_partition():
    bases, seg_off, eso, _ = synth.poa_batch(1, 200, depth=6, length=300, length_jitter=0.8, depth_jitter=4)
    cost = sharding.edge_costs(seg_off, eso)
    for world in (1, 2, 3, 8):
        sh = sharding.shard_edges(cost, world)
        allv = np.sort(np.concatenate(sh))
        assert np.array_equal(allv, np.arange(200))
        loads = np.array([cost[s].sum() for s in sh])
        assert loads.max() <= loads.mean() + cost.max() + 1e-9      # LPT bound


def test_take_shard_roundtrip():
    bases, seg_off, eso, _ = synth.poa_batch(2, 30, depth=4, length=100, length_jitter=0.5, depth_jitter=2)
    edges = np.array([3, 7, 29, 0])
    b, so, es = sharding.take_shard(bases, seg_off, eso, edges)
    assert len(es) == 5
    for k, e in enumerate(edges):
        for j in range(int(eso[e + 1] - eso[e])):
            s_old, s_new = int(eso[e]) + j, int(es[k]) + j
            assert bases[int(seg_off[s_old]): int(seg_off[s_old + 1])].tobytes() == b[int(so[s_new]): int(so[s_new + 1])].tobytes()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_ffi
    bases, seg_off, eso, _ = synth.poa_batch(7, 24, depth=4, length=120, length_jitter=0.4)
    sh = sharding.shard_edges(sharding.edge_costs(seg_off, eso), world)
    b, so, es = sharding.take_shard(bases, seg_off, eso, sh[rank])
    cons, off, _, _ = oracle_ffi.poa_batch(b, so, es)        # stand-in for the GPU call: this test is about the exchange
    got = sharding.all_gather_consensus(dist, torch.from_numpy(np.ascontiguousarray(cons)), off, torch.device("cpu"))
    # reassemble in global edge order and compare with the unsharded result
    full = [None] * 24
    for r in range(world):
        cb, co = got[r]
        for k, e in enumerate(sh[r]):
            full[e] = cb[int(co[k]): int(co[k + 1])].tobytes()
    rc, ro, _, _ = oracle_ffi.poa_batch(bases, seg_off, eso)
    ok = all(full[e] == rc[int(ro[e]): int(ro[e + 1])].tobytes() for e in range(24))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_all_gather_consensus_gloo_world2(oracle):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
