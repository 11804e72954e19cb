"""Batch packing on the caller side of the hot path.

The reference ships every batch's adjacency as a [batch][num_choice] nested Python list of small
int64 tensors, moves each one to the device separately (utils/data_utils.py:64-68) and offsets /
concatenates them per forward (modeling_qagnn.py:244-251): 2*bs*nc tiny H2D copies and bs*nc tiny
kernels.  `pack_adj` does the same offsetting and concatenation ONCE on the host into two pinned
tensors, so a batch needs two H2D copies.  The nested-list format keeps working everywhere.
"""
import os
import pickle
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class PackedAdj:
    """edge_index int64 [2, total_E] with per-graph node offsets already applied, edge_type int64 [total_E],
    graph_ptr int64 [n_graphs + 1] = first edge of each graph."""
    edge_index: torch.Tensor
    edge_type: torch.Tensor
    graph_ptr: torch.Tensor
    n_nodes: int

    def to(self, device, non_blocking=True):
        return PackedAdj(self.edge_index.to(device, non_blocking=non_blocking),
                         self.edge_type.to(device, non_blocking=non_blocking), self.graph_ptr, self.n_nodes)

    # LM_QAGNN.forward slices inputs positionally and calls .size() only on tensors before the last two
    def __iter__(self):
        return iter((self.edge_index, self.edge_type))


def pack_adj(edge_index_nested, edge_type_nested, n_nodes, pin=True):
    """[batch][num_choice] nested lists (load_sparse_adj_data_with_contextnode format,
    utils/data_utils.py:189-190) -> PackedAdj, equal to LM_QAGNN.batch_graph's output."""
    flat_ei = [e for row in edge_index_nested for e in row] if isinstance(edge_index_nested[0], (list, tuple)) else list(edge_index_nested)
    flat_et = [e for row in edge_type_nested for e in row] if isinstance(edge_type_nested[0], (list, tuple)) else list(edge_type_nested)
    counts = torch.tensor([e.size(1) for e in flat_ei], dtype=torch.long)
    graph_ptr = torch.zeros(len(flat_ei) + 1, dtype=torch.long)
    graph_ptr[1:] = torch.cumsum(counts, 0)
    total = int(graph_ptr[-1])
    edge_index = torch.empty(2, total, dtype=torch.long, pin_memory=pin and torch.cuda.is_available())
    edge_type = torch.empty(total, dtype=torch.long, pin_memory=pin and torch.cuda.is_available())
    if total:
        torch.cat(flat_ei, dim=1, out=edge_index)
        torch.cat(flat_et, dim=0, out=edge_type)
        edge_index += torch.repeat_interleave(torch.arange(len(flat_ei)) * n_nodes, counts).unsqueeze(0)
    return PackedAdj(edge_index, edge_type, graph_ptr, n_nodes)


def load_sparse_adj_data_with_contextnode(adj_pk_path, max_node_num, num_choice, args=None, use_cache=True,
                                          write_cache=True):
    """Same contract as the reference loader (utils/data_utils.py:79-197): reads the `*.graph.adj.pk` list of
    {'adj': scipy COO (half_n_rel*n x n), 'concepts', 'qmask', 'amask', 'cid2score'} records (or the
    `.loaded_cache` pickle next to it, same 8-item layout) and returns

        concept_ids [Q, nc, n] int64, node_type_ids [Q, nc, n] int64, node_scores [Q, nc, n, 1] fp32,
        adj_lengths [Q, nc] int64, (edge_index, edge_type) nested lists [Q][nc] of int64 [2, E_g] / [E_g]

    with the context node at position 0 (concept id 0, type 3, cid2score key -1), concept ids shifted by +1, pad
    id 1 / type 2, relation ids shifted by +2 with context->question / context->answer edges as relations 0 / 1, edges
    to truncated nodes dropped and the inverse edges appended with `rel + half_n_rel`.  The per-record work is
    vectorised (no per-node Python loops), which is what makes the reference's loader take minutes on CSQA."""
    cache_path = adj_pk_path + ".loaded_cache"
    if use_cache and os.path.exists(cache_path):
        with open(cache_path, "rb") as f:
            (adj_lengths_ori, concept_ids, node_type_ids, node_scores, adj_lengths, edge_index, edge_type,
             half_n_rel) = pickle.load(f)
    else:
        with open(adj_pk_path, "rb") as fin:
            records = pickle.load(fin)
        n_samples = len(records)
        edge_index, edge_type = [], []
        adj_lengths = np.zeros(n_samples, dtype=np.int64)
        adj_lengths_ori = np.zeros(n_samples, dtype=np.int64)
        concept_ids = np.ones((n_samples, max_node_num), dtype=np.int64)
        node_type_ids = np.full((n_samples, max_node_num), 2, dtype=np.int64)
        node_scores = np.zeros((n_samples, max_node_num, 1), dtype=np.float32)
        half_n_rel = 0
        for idx, rec in enumerate(records):
            adj, concepts = rec["adj"], np.asarray(rec["concepts"])
            qm, am = np.asarray(rec["qmask"], dtype=bool), np.asarray(rec["amask"], dtype=bool)
            cid2score = rec["cid2score"]
            if len(concepts) != len(set(concepts.tolist())):
                raise ValueError(f"record {idx}: duplicate concepts")
            qam = qm | am
            if not qam[0] or np.any(np.diff(qam.astype(np.int8)) > 0):
                raise ValueError(f"record {idx}: question/answer concepts must form a prefix")
            num_concept = min(len(concepts), max_node_num - 1) + 1
            adj_lengths_ori[idx] = len(concepts)
            adj_lengths[idx] = num_concept
            kept = concepts[:num_concept - 1]
            concept_ids[idx, 0] = 0
            concept_ids[idx, 1:num_concept] = kept + 1
            if cid2score is not None:
                node_scores[idx, 0, 0] = cid2score[-1]
                node_scores[idx, 1:num_concept, 0] = [cid2score[int(c)] for c in kept]
            node_type_ids[idx, 0] = 3
            types = node_type_ids[idx, 1:num_concept]
            types[qm[:num_concept - 1]] = 0
            types[am[:num_concept - 1]] = 1
            n_node = adj.shape[1]
            half = adj.shape[0] // n_node
            rel = adj.row.astype(np.int64) // n_node + 2
            src = adj.row.astype(np.int64) % n_node + 1
            tgt = adj.col.astype(np.int64) + 1
            # context node -> question / answer concepts (relations 0 / 1); the reference's `> num_concept` bound
            limit = min(len(qm), num_concept)
            q_to = np.nonzero(qm[:limit])[0] + 1
            a_to = np.nonzero(am[:limit])[0] + 1
            rel = np.concatenate([rel, np.zeros(len(q_to), np.int64), np.ones(len(a_to), np.int64)])
            src = np.concatenate([src, np.zeros(len(q_to) + len(a_to), np.int64)])
            tgt = np.concatenate([tgt, q_to, a_to])
            half_n_rel = half + 2
            keep = (src < max_node_num) & (tgt < max_node_num)
            rel, src, tgt = rel[keep], src[keep], tgt[keep]
            edge_index.append(torch.from_numpy(np.stack([np.concatenate([src, tgt]), np.concatenate([tgt, src])])))
            edge_type.append(torch.from_numpy(np.concatenate([rel, rel + half_n_rel])))
        adj_lengths_ori = torch.from_numpy(adj_lengths_ori)
        adj_lengths = torch.from_numpy(adj_lengths)
        concept_ids = torch.from_numpy(concept_ids)
        node_type_ids = torch.from_numpy(node_type_ids)
        node_scores = torch.from_numpy(node_scores)
        if write_cache:
            with open(cache_path, "wb") as f:
                pickle.dump([adj_lengths_ori, concept_ids, node_type_ids, node_scores, adj_lengths, edge_index, edge_type,
                             half_n_rel], f)
    edge_index = [list(edge_index[i:i + num_choice]) for i in range(0, len(edge_index), num_choice)]
    edge_type = [list(edge_type[i:i + num_choice]) for i in range(0, len(edge_type), num_choice)]
    concept_ids, node_type_ids, node_scores, adj_lengths = [x.view(-1, num_choice, *x.size()[1:]) for x in
                                                            (concept_ids, node_type_ids, node_scores, adj_lengths)]
    return concept_ids, node_type_ids, node_scores, adj_lengths, (edge_index, edge_type)


def synth_adj_pickle(path, n_records, seed=0, n_rel=17, max_nodes=400):
    """Writes a synthetic `*.graph.adj.pk` in the reference's record schema (utils/graph.py:331-338): n_g ~
    clip(lognormal(ln 120, 0.6), 8, max_nodes) concepts, ~2.5*n_g forward edges over `n_rel` relations, 1-8 question
    and 1-3 answer concepts first, scores ~ N(0,1) incl. key -1 for the context node (SURVEY.md §8d cfg 3)."""
    from scipy.sparse import coo_matrix
    rng = np.random.default_rng(seed)
    records = []
    for _ in range(n_records):
        n = int(np.clip(rng.lognormal(np.log(120), 0.6), 8, max_nodes))
        concepts = rng.choice(700000, size=n, replace=False).astype(np.int64)
        nq = int(rng.integers(1, 9)); na = int(rng.integers(1, 4))
        nq = min(nq, n - 1); na = min(na, n - nq)
        qmask = np.zeros(n, dtype=bool); qmask[:nq] = True
        amask = np.zeros(n, dtype=bool); amask[nq:nq + na] = True
        ne = int(2.5 * n)
        r, s, t = rng.integers(0, n_rel, ne), rng.integers(0, n, ne), rng.integers(0, n, ne)
        key = np.unique(np.stack([r * n + s, t]), axis=1)
        adj = coo_matrix((np.ones(key.shape[1], dtype=bool), (key[0], key[1])), shape=(n_rel * n, n))
        cid2score = {int(c): float(v) for c, v in zip(concepts, rng.standard_normal(n))}
        cid2score[-1] = float(rng.standard_normal())
        records.append({"adj": adj, "concepts": concepts, "qmask": qmask, "amask": amask, "cid2score": cid2score})
    with open(path, "wb") as f:
        pickle.dump(records, f)
    return records
