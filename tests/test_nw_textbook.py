"""A second opinion on the POA dynamic programme that does not go through the restated SPOA: the graph-NW recurrence written
down from its textbook definition in numpy (global alignment, linear gaps, sequence-to-DAG as in Lee, Grasso & Sharlow 2002:
a cell takes the best of its predecessors' diagonal / vertical moves and its own horizontal move), evaluated on the graph the
implementation reports, and the score of the alignment it returns recomputed move by move.

Checked here: (1) the whole score matrix equals the textbook recurrence, (2) the returned alignment is a walk through the graph
from the source side to a sink that consumes the whole sequence, and (3) its score equals the optimum of the matrix - so the
traceback returns AN optimal alignment whatever the tie-breaking is. What stays pinned only on the restatement is SPOA's choice
among equal optima, its graph update and its consensus. The reference reaches this code through
spoa::AlignmentEngine::align_sequence_with_graph (Assemble.cpp:539, kNW, 5/-4/-8)."""
import numpy as np
import pytest

import synth


def textbook_graph_nw(dbg, seq, m, x, g):
    V, L = dbg["V"], dbg["L"]
    r2n = dbg["rank2node"].astype(np.int64)
    n2r = np.empty(V, dtype=np.int64)
    n2r[r2n] = np.arange(V)
    poff, pnode, code = dbg["pred_off"], dbg["pred_node"], dbg["code"]
    j = np.arange(L + 1, dtype=np.int64)
    H = np.zeros((V + 1, L + 1), dtype=np.int64)
    H[0] = j * g
    for r in range(V):
        rows = [int(n2r[p]) + 1 for p in pnode[poff[r]: poff[r + 1]]] or [0]      # no in-edge: the virtual start row
        assert all(p <= r for p in rows), "ranks are not a topological order"
        sub = np.where(seq == code[r2n[r]], m, x).astype(np.int64)
        best = np.full(L + 1, np.iinfo(np.int64).min // 2, dtype=np.int64)
        for p in rows:
            best[0] = max(best[0], H[p][0] + g)
            best[1:] = np.maximum(best[1:], np.maximum(H[p][:-1] + sub, H[p][1:] + g))
        # horizontal moves: H[i][j] = max(best[j], H[i][j-1] + g)  <=>  a running maximum of best[j] - j g
        H[r + 1] = np.maximum.accumulate(best - j * g) + j * g
    return H, n2r


def check_alignment(dbg, H, n2r, seq, m, x, g):
    V, L = dbg["V"], dbg["L"]
    poff, pnode, code = dbg["pred_off"], dbg["pred_node"], dbg["code"]
    an, ap = dbg["aln_node"], dbg["aln_pos"]
    has_out = np.zeros(V, dtype=bool)
    has_out[pnode] = True
    sinks = [int(n2r[v]) + 1 for v in range(V) if not has_out[v]]
    optimum = max(int(H[i][L]) for i in sinks)
    score, last_node, next_pos = 0, -1, 0
    for v, p in zip(an.tolist(), ap.tolist()):
        assert v >= 0 or p >= 0
        if p >= 0:
            assert p == next_pos, "sequence positions are not consecutive"
            next_pos += 1
        if v >= 0:
            preds = set(pnode[poff[n2r[v]]: poff[n2r[v] + 1]].tolist())
            assert (last_node in preds) if last_node >= 0 else (len(preds) == 0), "the alignment does not follow the graph's edges"
            last_node = v
        score += g if (v < 0 or p < 0) else (m if code[v] == seq[p] else x)
    assert next_pos == L, "the alignment does not consume the whole sequence"
    assert last_node >= 0 and not has_out[last_node], "the alignment does not end in a sink"
    assert score == optimum, f"alignment scores {score}, the matrix optimum is {optimum}"


CASES = [  # (seed, depth, length, jitter, n_prior)
    (41, 2, 300, 0.0, 1),        # chain graph: plain Needleman-Wunsch
    (42, 6, 400, 0.2, 3),
    (43, 9, 700, 0.1, 8),        # wide graph, two stripes
    (44, 5, 1300, 0.0, 4),       # three stripes
]


def run_case(dbg_fn, case, scores=(5, -4, -8), **kw):
    seed, depth, length, jitter, n_prior = case
    bases, seg_off, _, _ = synth.poa_batch(seed, 1, depth=depth, length=length, length_jitter=jitter)
    dbg = dbg_fn(bases, seg_off, n_prior, *scores, **kw)
    seq = bases[int(seg_off[n_prior]): int(seg_off[n_prior + 1])]
    assert dbg["L"] == len(seq)
    H, n2r = textbook_graph_nw(dbg, seq, *scores)
    assert np.array_equal(dbg["H"].astype(np.int64), H), "score matrix differs from the textbook recurrence"
    check_alignment(dbg, H, n2r, seq, *scores)


@pytest.mark.parametrize("case", CASES)
def test_oracle_against_textbook_recurrence(oracle, case):
    run_case(oracle.poa_debug, case)


def test_oracle_other_scores(oracle):
    run_case(oracle.poa_debug, CASES[1], scores=(3, -5, -4))
    run_case(oracle.poa_debug, CASES[1], scores=(1, -1, -1))


@pytest.mark.gpu
@pytest.mark.parametrize("force_i32", [0, 1, 2])
@pytest.mark.parametrize("case", CASES)
def test_kernels_against_textbook_recurrence(ctx, case, force_i32):
    """The CUDA fill and traceback in each cell encoding (0 = picked, 1 = int32, 2 = row-relative int16), no oracle involved."""
    run_case(ctx.poa_debug, case, force_i32=force_i32)


@pytest.mark.gpu
def test_kernels_other_scores(ctx):
    for scores in ((3, -5, -4), (1, -1, -1)):
        for force in (0, 2):
            run_case(ctx.poa_debug, CASES[1], scores=scores, force_i32=force)
