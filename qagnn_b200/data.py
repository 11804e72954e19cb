"""Batch packing on the caller side of the hot path.

The reference ships every batch's adjacency as a [batch][num_choice] nested Python list of small
int64 tensors, moves each one to the device separately (utils/data_utils.py:64-68) and offsets /
concatenates them per forward (modeling_qagnn.py:244-251): 2*bs*nc tiny H2D copies and bs*nc tiny
kernels.  `pack_adj` does the same offsetting and concatenation ONCE on the host into two pinned
tensors, so a batch needs two H2D copies.  The nested-list format keeps working everywhere.
"""
from dataclasses import dataclass

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
