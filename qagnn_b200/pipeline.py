"""Host<->device streaming around the hot path (caller side).

`StreamedRunner` feeds batches that live in pinned host memory through `QAGNN_Message_Passing.forward` with the
copies on their own CUDA streams: while batch i computes, batch i+1 is uploaded and the result of batch i-1 is
downloaded (PCIe is full duplex), each batch in one of `depth` device buffer sets.  Every batch still pays its own
H2D and D2H; they just stop serialising with the kernels.  With `module.use_cuda_graph = True` each buffer set gets
its own captured graph (static addresses).
"""
import torch


class StreamedRunner:
    FIELDS = ("H", "edge_index", "edge_type", "node_type", "node_score")

    def __init__(self, module, example, device, depth=2):
        """`example`: dict of pinned host tensors (keys FIELDS) giving the shapes/dtypes of every batch."""
        self.module, self.device, self.depth = module, device, depth
        self.h2d, self.compute, self.d2h = (torch.cuda.Stream(device=device) for _ in range(3))
        self.dev = [{k: torch.empty_like(example[k], device=device) for k in self.FIELDS} for _ in range(depth)]
        B, n, D = example["H"].shape
        self.host_out = [torch.empty(B, n, D, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.ev_up = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.ev_down = [torch.cuda.Event() for _ in range(depth)]
        self.count = 0
        self.dev_out = [None] * depth   # device result of the last batch submitted to each slot
        cur = torch.cuda.current_stream(device)
        for s in (self.h2d, self.compute, self.d2h):
            s.wait_stream(cur)

    def submit(self, host_batch):
        """Enqueues one batch (dict of pinned host tensors); returns the slot whose `host_out` will hold the result
        once `ev_down[slot]` has completed (or after `drain()`)."""
        s = self.count % self.depth
        self.count += 1
        self.h2d.wait_event(self.ev_done[s])        # the previous batch in this slot no longer reads the inputs
        with torch.cuda.stream(self.h2d):
            for k in self.FIELDS:
                self.dev[s][k].copy_(host_batch[k], non_blocking=True)
            self.ev_up[s].record(self.h2d)
        self.compute.wait_event(self.ev_up[s])
        self.compute.wait_event(self.ev_down[s])    # the slot's output buffer has been downloaded
        with torch.cuda.stream(self.compute):
            d = self.dev[s]
            out = self.module(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"])
            self.dev_out[s] = out
            self.ev_done[s].record(self.compute)
        self.d2h.wait_event(self.ev_done[s])
        with torch.cuda.stream(self.d2h):
            self.host_out[s].copy_(out, non_blocking=True)
            self.ev_down[s].record(self.d2h)
        return s

    def drain(self):
        cur = torch.cuda.current_stream(self.device)
        for s in (self.h2d, self.compute, self.d2h):
            cur.wait_stream(s)
