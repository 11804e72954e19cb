"""Pins the integer side of the oracle (oracle.qagnn_oracle.graph_prep_oracle) against a naive per-edge Python
restatement of the reference's bookkeeping on small cases: the one-hot edge feature of
modeling_qagnn.py:419-432 (real edges: [onehot(etype, R+1) | onehot(type[src], T) | onehot(type[tgt], T)];
the N appended self loops: etype R, both types the node's own), edge_index' with the loops appended after the real
edges (:436-438), and the out-degree that rescales alpha (:476-479)."""
import numpy as np
import pytest
import torch

from oracle import qagnn_oracle as O


def _naive(edge_index, edge_type, node_type, T, R):
    ei, et, nt = edge_index.tolist(), edge_type.tolist(), node_type.reshape(-1).tolist()
    N, E = len(nt), len(et)
    feats, src, tgt = [], [], []
    for e in range(E):
        f = [0] * (R + 1 + 2 * T)
        f[et[e]] = 1
        f[R + 1 + nt[ei[0][e]]] = 1
        f[R + 1 + T + nt[ei[1][e]]] = 1
        feats.append(f); src.append(ei[0][e]); tgt.append(ei[1][e])
    for v in range(N):
        f = [0] * (R + 1 + 2 * T)
        f[R] = 1
        f[R + 1 + nt[v]] = 1
        f[R + 1 + T + nt[v]] = 1
        feats.append(f); src.append(v); tgt.append(v)
    outdeg = [0] * N
    for s in src:
        outdeg[s] += 1
    by_src = sorted(range(E + N), key=lambda e: (src[e], e))  # stable order by source
    by_tgt = sorted(range(E + N), key=lambda e: (tgt[e], e))
    return feats, src, tgt, outdeg, by_src, by_tgt


def _feature_of_combo(c, T, R):
    f = [0] * (R + 1 + 2 * T)
    if c >= R * T * T:  # self loop of a node of type c - R*T*T
        t = c - R * T * T
        f[R] = 1; f[R + 1 + t] = 1; f[R + 1 + T + t] = 1
    else:
        et, ts, tt = c // (T * T), (c // T) % T, c % T
        f[et] = 1; f[R + 1 + ts] = 1; f[R + 1 + T + tt] = 1
    return f


@pytest.mark.parametrize("seed,B,n,e,R", [(0, 3, 7, 20, 38), (1, 1, 1, 0, 38), (2, 2, 5, 40, 6), (3, 4, 9, 1, 17)])
def test_prep_oracle_matches_naive_restatement(seed, B, n, e, R):
    T = 4
    g = torch.Generator().manual_seed(seed)
    node_type = torch.randint(0, T, (B, n), generator=g)
    ei = torch.randint(0, n, (B, 2, e), generator=g) + (torch.arange(B) * n).view(B, 1, 1)
    edge_index = ei.permute(1, 0, 2).reshape(2, B * e).contiguous()
    edge_type = torch.randint(0, R, (B * e,), generator=g)
    ref = O.graph_prep_oracle(edge_index, edge_type, node_type, T, R)
    feats, src, tgt, outdeg, by_src, by_tgt = _naive(edge_index, edge_type, node_type, T, R)
    assert ref["src"].tolist() == src and ref["tgt"].tolist() == tgt
    assert ref["outdeg"].tolist() == outdeg and min(outdeg) >= 1  # the self loop
    assert ref["perm_src"].tolist() == by_src and ref["perm_tgt"].tolist() == by_tgt
    assert ref["rowptr_src"].tolist() == np.concatenate([[0], np.cumsum(outdeg)]).tolist()
    C = R * T * T + T
    assert 0 <= ref["combo"].min() and ref["combo"].max() < C
    for e_id, c in enumerate(ref["combo"].tolist()):
        assert _feature_of_combo(c, T, R) == feats[e_id], e_id
    # distinct features <-> distinct combos (the folded tables are indexed by it)
    seen = {}
    for f, c in zip(feats, ref["combo"].tolist()):
        assert seen.setdefault(tuple(f), c) == c
