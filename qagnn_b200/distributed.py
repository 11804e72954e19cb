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


def all_gather_rows(local, world_size, group=None, equal_shards=True, out=None):
    """The path's ONE collective: all-gather of the ranks' row blocks [B_r, F] -> [sum B_r, F].

    equal_shards=True (how `shard_bounds` cuts a batch whose question count divides by the world size, and cfg4 of
    BASELINE.json): a single `all_gather_into_tensor`, no size exchange, no host synchronisation — capturable in a CUDA
    graph.  `out` optionally names the [world_size * B_r, F] result buffer (static address for graph replay).
    equal_shards=False: blocks may differ in length (a short last shard): sizes are exchanged first, then padded blocks."""
    if world_size == 1:
        return local
    if equal_shards:
        if out is None:
            out = local.new_empty((world_size * local.size(0),) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
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
                            group=None, equal_shards=True):
    """QAGNN.forward on this rank's shard + all-gather of the pooled features + the answer MLP on the full batch
    (modeling_qagnn.py:172-188).  Returns (logits of ALL graphs [B_total, 1], local pool_attn)."""
    concat, pool_attn = decoder.pooled_features(sent_vecs, concept_ids, node_type_ids, node_scores, adj_lengths, adj)
    full = all_gather_rows(concat, world_size, group, equal_shards)
    return decoder.fc(full), pool_attn


def bind_to_gpu_numa_node(device_index):
    """Pins this process (and the pinned host buffers it allocates afterwards, by first touch) to the CPUs of the NUMA node
    the GPU hangs off.  8 ranks streaming ~110 MB per step through pinned memory otherwise cross the socket interconnect
    for half of the GPUs (SCALE_r01: e2e efficiency 0.74 at N=8 with device-side efficiency 0.97).  Returns the node or None."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001 - affinity is an optimisation, never a failure
        return None
