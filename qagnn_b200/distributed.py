"""Data-parallel use of the path across the GPUs of one box.

Sub-graphs are independent (LM_QAGNN.batch_graph only offsets node ids, modeling_qagnn.py:244-251), weights are
tiny and replicated, so the (question, all-choices) groups are sharded across ranks with no collective inside the k
GNN layers.  The single exchange is an all-gather of the pooled features right before the answer MLP
(modeling_qagnn.py:187-188): every rank ends up with the logits of the whole batch.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_questions, rank, world_size):
    """Contiguous block of whole questions for `rank` (all choices of a question stay together)."""
    per = (n_questions + world_size - 1) // world_size
    lo = min(rank * per, n_questions)
    return lo, min(lo + per, n_questions)


def shard_batch(inputs, rank, world_size):
    """Slices an LM_QAGNN-style positional input tuple (tensors with a leading batch dim + the two nested
    [batch][num_choice] adjacency lists at the end) down to this rank's questions."""
    *tensors, edge_index, edge_type = inputs
    lo, hi = shard_bounds(tensors[0].size(0), rank, world_size)
    return tuple(t[lo:hi] for t in tensors) + (edge_index[lo:hi], edge_type[lo:hi]), (lo, hi)


def all_gather_rows(local, world_size, group=None):
    """All-gather of row blocks that may differ in length by rank (the last shard can be short)."""
    if world_size == 1:
        return local
    n = torch.tensor([local.size(0)], device=local.device, dtype=torch.long)
    sizes = [torch.zeros_like(n) for _ in range(world_size)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    mx = max(sizes)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.size(0)] = local
    out = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def decoder_forward_sharded(decoder, sent_vecs, concept_ids, node_type_ids, node_scores, adj_lengths, adj, world_size,
                            group=None):
    """QAGNN.forward on this rank's shard + all-gather of the pooled features + the answer MLP on the full batch.
    Returns (logits of ALL graphs [B_total, 1], local pool_attn)."""
    concat, pool_attn = decoder.pooled_features(sent_vecs, concept_ids, node_type_ids, node_scores, adj_lengths, adj)
    full = all_gather_rows(concat, world_size, group)
    return decoder.fc(full), pool_attn
