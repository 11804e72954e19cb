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
    graph_ptr int64 [n_graphs + 1] = first edge of each graph.  When built by FlatAdjCache.pack the two tensors are views
    of ONE [3, total_E] buffer (`buf`), so the whole adjacency of a batch moves to the device in a single copy."""
    edge_index: torch.Tensor
    edge_type: torch.Tensor
    graph_ptr: torch.Tensor
    n_nodes: int
    buf: torch.Tensor = None

    graph_ptr_dev: torch.Tensor = None   # graph_ptr on the device of edge_index (packed graph prep, one launch per batch)
    max_edges: int = -1                  # max edges of one sub-graph (host value; -1 = derive from graph_ptr)

    def to(self, device, non_blocking=True):
        gp = self.graph_ptr
        if torch.cuda.is_available() and torch.device(device).type == "cuda" and not gp.is_cuda and not gp.is_pinned():
            gp = gp.pin_memory()
        gpd = gp.to(device, non_blocking=non_blocking)
        me = self.max_edges if self.max_edges >= 0 else (int((self.graph_ptr[1:] - self.graph_ptr[:-1]).max()) if self.graph_ptr.numel() > 1 else 0)
        if self.buf is not None:
            b = self.buf.to(device, non_blocking=non_blocking)
            return PackedAdj(b[:2], b[2], self.graph_ptr, self.n_nodes, b, gpd, me)
        return PackedAdj(self.edge_index.to(device, non_blocking=non_blocking),
                         self.edge_type.to(device, non_blocking=non_blocking), self.graph_ptr, self.n_nodes, None, gpd, me)

    # LM_QAGNN.forward slices inputs positionally and calls .size() only on tensors before the last two; everywhere an
    # (edge_index, edge_type) pair is expected a PackedAdj unpacks / indexes as that pair
    def __iter__(self):
        return iter((self.edge_index, self.edge_type))

    def __getitem__(self, i):
        return (self.edge_index, self.edge_type)[i]

    def __len__(self):
        return 2


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


# ------------------------------------------------------------------------------------------------------------------
# Flat adjacency cache + pre-packed batch generator
# ------------------------------------------------------------------------------------------------------------------
class FlatAdjCache:
    """All sub-graphs of a split as three flat arrays + a row pointer (a CSR over graphs), instead of the reference's
    [n_samples][num_choice] nested lists of small tensors (utils/data_utils.py:174-190):

        src, tgt int32 [total_E] (LOCAL node ids), etype int16 [total_E], graph_ptr int64 [n_graphs + 1]

    Written once next to the reference's `.loaded_cache` as `<adj_pk_path>.flat_cache.npz` and memory-mapped afterwards;
    `pack(graph_ids)` cuts a batch out of it with three vectorised gathers into ONE pinned [3, E] int64 buffer."""

    def __init__(self, src, tgt, etype, graph_ptr, num_choice, n_nodes):
        self.src, self.tgt, self.etype, self.graph_ptr = src, tgt, etype, graph_ptr
        self.num_choice, self.n_nodes = int(num_choice), int(n_nodes)

    @classmethod
    def from_nested(cls, edge_index_nested, edge_type_nested, n_nodes):
        flat_ei = [e for row in edge_index_nested for e in row]
        flat_et = [e for row in edge_type_nested for e in row]
        counts = np.array([e.size(1) for e in flat_ei], dtype=np.int64)
        ptr = np.zeros(len(flat_ei) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum(counts)
        if int(ptr[-1]):
            ei = torch.cat(flat_ei, dim=1).numpy()
            et = torch.cat(flat_et, dim=0).numpy()
        else:
            ei, et = np.zeros((2, 0), np.int64), np.zeros((0,), np.int64)
        return cls(ei[0].astype(np.int32), ei[1].astype(np.int32), et.astype(np.int16), ptr, len(edge_index_nested[0]), n_nodes)

    def save(self, path):
        np.savez(path, src=self.src, tgt=self.tgt, etype=self.etype, graph_ptr=self.graph_ptr,
                 meta=np.array([self.num_choice, self.n_nodes], dtype=np.int64))

    @classmethod
    def load(cls, path):
        z = np.load(path, mmap_mode="r")
        meta = z["meta"]
        return cls(z["src"], z["tgt"], z["etype"], z["graph_ptr"], int(meta[0]), int(meta[1]))

    def n_graphs(self):
        return len(self.graph_ptr) - 1

    def n_questions(self):
        return self.n_graphs() // self.num_choice

    def __getitem__(self, i):
        """What LM_QAGNN_DataLoader does with `adj_data` besides handing it to the batch generator
        (modeling_qagnn.py:281-287,307-308): `len(adj_data[0])` must be the number of questions, and `adj_data[:n]` is taken
        when a split is subsampled.  An int gives a sized stand-in for the nested list, a slice the cache of those questions."""
        if isinstance(i, slice):
            a, b, step = i.indices(self.n_questions())
            if step != 1:
                raise IndexError("FlatAdjCache supports contiguous question ranges only")
            b = max(a, b)
            g0, g1 = a * self.num_choice, b * self.num_choice
            e0, e1 = int(self.graph_ptr[g0]), int(self.graph_ptr[g1])
            return FlatAdjCache(self.src[e0:e1], self.tgt[e0:e1], self.etype[e0:e1],
                                np.asarray(self.graph_ptr[g0:g1 + 1]) - e0, self.num_choice, self.n_nodes)
        if i in (0, 1, -1, -2):
            return range(self.n_questions())
        raise IndexError("FlatAdjCache stands in for the (edge_index, edge_type) pair: index 0 or 1")

    def pack(self, question_indexes, pin=True):
        """PackedAdj of the questions `question_indexes` (all their choices, in order) — equal to
        LM_QAGNN.batch_graph applied to the corresponding nested-list slice (modeling_qagnn.py:244-251)."""
        q = np.asarray(question_indexes, dtype=np.int64).reshape(-1)
        gids = (q[:, None] * self.num_choice + np.arange(self.num_choice)[None, :]).reshape(-1)
        beg, end = self.graph_ptr[gids], self.graph_ptr[gids + 1]
        counts = end - beg
        ptr = np.zeros(len(gids) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum(counts)
        total = int(ptr[-1])
        # flat positions of every edge of the batch: beg[g] + (0 .. counts[g]-1)
        pos = np.repeat(beg - ptr[:-1], counts) + np.arange(total, dtype=np.int64)
        off = np.repeat(np.arange(len(gids), dtype=np.int64) * self.n_nodes, counts)
        buf = torch.empty(3, total, dtype=torch.long, pin_memory=pin and torch.cuda.is_available())
        b = buf.numpy()
        np.add(self.src[pos], off, out=b[0], casting="unsafe")
        np.add(self.tgt[pos], off, out=b[1], casting="unsafe")
        b[2] = self.etype[pos]
        return PackedAdj(buf[:2], buf[2], torch.from_numpy(ptr), self.n_nodes, buf)


def load_flat_adj_cache(adj_pk_path, max_node_num, num_choice, args=None):
    """(concept_ids, node_type_ids, node_scores, adj_lengths, FlatAdjCache): the loader above with the adjacency as a flat
    cache; `<adj_pk_path>.flat_cache.npz` is written on first use and memory-mapped on later ones."""
    flat_path = adj_pk_path + ".flat_cache.npz"
    cids, ntypes, scores, lens, (ei, et) = load_sparse_adj_data_with_contextnode(adj_pk_path, max_node_num, num_choice, args)
    if os.path.exists(flat_path):
        flat = FlatAdjCache.load(flat_path)
        if flat.n_graphs() == cids.size(0) * num_choice and flat.n_nodes == max_node_num and flat.num_choice == num_choice:
            return cids, ntypes, scores, lens, flat
    flat = FlatAdjCache.from_nested(ei, et, max_node_num)
    flat.save(flat_path)
    return cids, ntypes, scores, lens, flat


class PackedAdjBatchGenerator:
    """Drop-in for the reference's MultiGPUSparseAdjDataBatchGenerator (utils/data_utils.py:17-76): same constructor, same
    batch tuple  (qids, labels, *tensors0, *lists0, *tensors1, *lists1, edge_index, edge_type),  same partial-batch
    options — but `adj_data` may be a FlatAdjCache, in which case the adjacency of a batch is packed on the host into one
    pinned buffer and moved with ONE non-blocking copy; the `edge_index` slot then holds a PackedAdj (which
    LM_QAGNN.forward accepts) and the `edge_type` slot its edge_type view.  Tensors are staged through pinned memory as
    well, so every copy of a batch is asynchronous.  With nested-list `adj_data` it behaves exactly like the reference."""

    def __init__(self, args, mode, device0, device1, batch_size, indexes, qids, labels, tensors0=[], lists0=[], tensors1=[],
                 lists1=[], adj_data=None, prefetch=0):
        """`prefetch` > 0 (not in the reference): a worker thread slices / packs / pins up to that many batches ahead on the
        host while the consumer trains on the current one; the device copies are still issued by the consuming thread (on
        ITS current stream), in order, so the batches and their order are exactly those of prefetch=0."""
        self.args, self.mode = args, mode
        self.device0, self.device1 = device0, device1
        self.batch_size, self.indexes, self.qids, self.labels = batch_size, indexes, qids, labels
        self.tensors0, self.lists0, self.tensors1, self.lists1 = tensors0, lists0, tensors1, lists1
        self.adj_data = adj_data
        self.prefetch = int(prefetch)

    def __len__(self):
        return (self.indexes.size(0) - 1) // self.batch_size + 1

    @staticmethod
    def _pin(obj, device):
        """Host side of a copy: page-locks a tensor bound for a CUDA device (nested lists element-wise)."""
        if isinstance(obj, (tuple, list)):
            return [PackedAdjBatchGenerator._pin(item, device) for item in obj]
        if isinstance(obj, PackedAdj):
            return obj  # FlatAdjCache.pack already wrote it into pinned memory
        if torch.cuda.is_available() and torch.device(device).type == "cuda" and not obj.is_pinned():
            return obj.pin_memory()
        return obj

    def _to_device(self, obj, device):
        if isinstance(obj, (tuple, list)):
            return [self._to_device(item, device) for item in obj]
        if isinstance(obj, PackedAdj):
            return obj.to(device, non_blocking=True)
        return self._pin(obj, device).to(device, non_blocking=True)

    def _host_batch(self, batch_indexes):
        """Everything of one batch that needs no device: slicing, adjacency packing, pinning."""
        d0, d1 = self.device0, self.device1
        qids = [self.qids[idx] for idx in batch_indexes]
        labels = self._pin(self.labels[batch_indexes], d1)
        tensors0 = [self._pin(x[batch_indexes], d0) for x in self.tensors0]
        tensors1 = [self._pin(x[batch_indexes], d1) for x in self.tensors1]
        lists0 = [self._pin([x[i] for i in batch_indexes], d0) for x in self.lists0]
        lists1 = [self._pin([x[i] for i in batch_indexes], d1) for x in self.lists1]
        if isinstance(self.adj_data, FlatAdjCache):
            adj = self.adj_data.pack(batch_indexes.numpy())
        else:
            edge_index_all, edge_type_all = self.adj_data
            adj = (self._pin([edge_index_all[i] for i in batch_indexes], d1),
                   self._pin([edge_type_all[i] for i in batch_indexes], d1))
        return qids, labels, tensors0, lists0, tensors1, lists1, adj

    def _ship(self, host):
        qids, labels, tensors0, lists0, tensors1, lists1, adj = host
        d0, d1 = self.device0, self.device1
        if isinstance(adj, PackedAdj):
            packed = self._to_device(adj, d1)
            edge_index, edge_type = packed, packed.edge_type
        else:
            edge_index, edge_type = self._to_device(adj[0], d1), self._to_device(adj[1], d1)
        return tuple([qids, self._to_device(labels, d1), *[self._to_device(x, d0) for x in tensors0],
                      *[self._to_device(x, d0) for x in lists0], *[self._to_device(x, d1) for x in tensors1],
                      *[self._to_device(x, d1) for x in lists1], edge_index, edge_type])

    def _batch_ranges(self):
        bs = self.batch_size
        n = self.indexes.size(0)
        if self.mode == "train" and getattr(self.args, "drop_partial_batch", False):
            n = (n // bs) * bs
        elif self.mode == "train" and getattr(self.args, "fill_partial_batch", False):
            remain = n % bs
            if remain > 0:
                extra = np.random.choice(self.indexes[:-remain], size=(bs - remain), replace=False)
                self.indexes = torch.cat([self.indexes, torch.tensor(extra)])
                n = self.indexes.size(0)
        return [self.indexes[a:min(n, a + bs)] for a in range(0, n, bs)]

    def __iter__(self):
        batches = self._batch_ranges()
        if self.prefetch <= 0:
            for batch_indexes in batches:
                yield self._ship(self._host_batch(batch_indexes))
            return
        import queue
        import threading
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def put(item):  # a consumer that stops iterating early must not leave the worker blocked on a full queue
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    pass
            return False

        def work():
            try:
                for batch_indexes in batches:
                    if not put(("batch", self._host_batch(batch_indexes))):
                        return
                put(("end", None))
            except BaseException as e:  # surfaces in the consuming thread
                put(("error", e))

        t = threading.Thread(target=work, name="qagnn-batch-prefetch", daemon=True)
        t.start()
        try:
            while True:
                kind, item = q.get()
                if kind == "end":
                    break
                if kind == "error":
                    raise item
                yield self._ship(item)
        finally:
            stop.set()
            t.join()
