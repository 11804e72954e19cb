"""Drop-in modules for the reference's modeling/modeling_qagnn.py hot path.

Same class names, constructor arguments, forward signatures, return values and state_dict key
names as the reference, but `GATConvE.forward` and
`QAGNN_Message_Passing.forward` run entirely in the hand-written sm_100a kernels of
libqagnn_b200.so, reached through the C ABI in include/qagnn_b200.h.  The nn.Modules below only
own parameters and device buffers (PyTorch = memory + streams).

    GATConvE                 <- modeling/modeling_qagnn.py:380-484
    QAGNN_Message_Passing    <- modeling/modeling_qagnn.py:7-95
    QAGNN                    <- modeling/modeling_qagnn.py:99-189   (caller side, plain PyTorch)
    LM_QAGNN                 <- modeling/modeling_qagnn.py:192-251  (caller side, plain PyTorch)

Modes: `.eval()` = the fused inference forward (dropout = identity, BatchNorm running statistics folded into the
weights) — the reference's `evaluate_accuracy` path (qagnn.py:30-38); `.train()` = dropout, BatchNorm batch statistics
and autograd through the CUDA message-passing forward / backward kernels (qagnn_b200/training.py; qagnn.py:249-278).
CPU tensors raise: there is no CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .layers import GELU, MLP, CustomizedEmbedding, MultiheadAttPoolLayer


def make_one_hot(labels, C_):
    """int64 [M] -> float32 one-hot [M, C] on labels.device (modeling_qagnn.py:352-367)."""
    out = torch.zeros(labels.size(0), C_, dtype=torch.float32, device=labels.device)
    return out.scatter_(1, labels.unsqueeze(1), 1.0)


def _version_key(tensors):
    return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)


_warned_grad = set()


def _warn_if_grad_wanted(what, *tensors):
    """The eval-mode forward runs outside autograd (its outputs carry no grad_fn).  Say so once if a caller seems to
    expect gradients; `.train()` selects the differentiable path (qagnn_b200/training.py)."""
    if torch.is_grad_enabled() and what not in _warned_grad and any(t is not None and t.requires_grad for t in tensors):
        import warnings
        _warned_grad.add(what)
        warnings.warn(f"qagnn_b200.{what}: eval-mode forward is not differentiable (inputs require grad but the output "
                      f"has no grad_fn); call .train() for the autograd path")


class _DeviceBlob:
    """A grow-only uint8 device buffer (workspace the C ABI asks the caller to own).  `on_replace` runs before the
    buffer is dropped for a larger one: captured CUDA graphs hold its raw address and must be discarded first."""

    def __init__(self, on_replace=None):
        self.buf = None
        self.on_replace = on_replace

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            if self.buf is not None and self.on_replace is not None:
                self.on_replace()
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self.buf


class GraphPrep:
    """Owns the opaque graph-prep workspace of one batched graph (qagnn_graph_prep)."""

    def __init__(self, edge_index, edge_type, node_type, n_ntype, n_etype, n_per_graph=0, validate=True,
                 allow_general_fallback=False, graph_ptr=None, max_edges=-1):
        """graph_ptr (device int64 [n_graphs + 1]) + max_edges (host int): the batch is PACKED graph by graph — the whole
        prep is then one kernel launch (qagnn_graph_prep_packed) instead of twelve."""
        lib = _lib.load()
        ei = _lib.i64c(edge_index, "edge_index")
        et = _lib.i64c(edge_type, "edge_type")
        nt = _lib.i64c(node_type.reshape(-1), "node_type")
        if ei.dim() != 2 or ei.size(0) != 2 or et.dim() != 1 or et.numel() != ei.size(1):
            raise ValueError("edge_index must be [2,E] and edge_type [E]")
        self.N, self.E = nt.numel(), ei.size(1)
        self.T, self.R, self.n_per_graph = n_ntype, n_etype, int(n_per_graph)
        self.device = nt.device
        self.layout = _lib.PrepLayout()
        _lib.check(lib.qagnn_graph_prep_layout(self.N, self.E, C.byref(self.layout)), "qagnn_graph_prep_layout")
        self.buf = torch.empty(self.layout.total_bytes, dtype=torch.uint8, device=self.device)
        self._combo_order = None

        packed = graph_ptr is not None and max_edges >= 0 and self.n_per_graph > 0
        if packed:
            gp = _lib.i64c(graph_ptr, "graph_ptr")
            if gp.numel() != self.N // max(self.n_per_graph, 1) + 1:
                raise ValueError("graph_ptr must have one entry per sub-graph plus one")

        def run():
            shape = _lib.Shape(self.N, self.E, 4, 1, n_ntype, n_etype, 0, self.n_per_graph)
            with torch.cuda.device(self.device):
                if packed and self.n_per_graph > 0:
                    st_ = lib.qagnn_graph_prep_packed(_lib.ptr(ei) if self.E else None, _lib.ptr(et) if self.E else None, _lib.ptr(nt),
                                                      _lib.ptr(gp), int(max_edges), C.byref(shape), _lib.ptr(self.buf), self.buf.numel(),
                                                      int(bool(validate)), _lib.stream_ptr(self.device))
                    if st_ != -5:  # QAGNN_ERR_UNSUPPORTED: sub-graphs too large for one CTA -> the general pipeline below
                        return st_
                return lib.qagnn_graph_prep(_lib.ptr(ei) if self.E else None, _lib.ptr(et) if self.E else None, _lib.ptr(nt),
                                            C.byref(shape), _lib.ptr(self.buf), self.buf.numel(), int(bool(validate)),
                                            _lib.stream_ptr(self.device))
        st = run()
        if st == -3 and self.n_per_graph > 0 and allow_general_fallback and int(self.array("status")[0].item()) == 8:
            packed = False
            # the only complaint is an edge that crosses a sub-graph boundary: n_per_graph is a layout HINT for the
            # shared-memory-tiled kernels, the reference accepts any batched edge_index -> use the general CSR kernels
            self.n_per_graph = 0
            st = run()
        _lib.check(st, "qagnn_graph_prep")
        self._keep = (ei, et, nt, graph_ptr)

    def combo_order(self):
        """int32 [E+N]: by-source edge positions stably sorted by combo (for the backward's edge-table gradient)."""
        if self._combo_order is None:
            self._combo_order = torch.argsort(self.array("csr_src_combo"), stable=True).to(torch.int32)
        return self._combo_order

    def combo_count(self):
        """int64 [C]: number of edges (self loops included) per combo index."""
        return torch.bincount(self.array("combo").long(), minlength=self.R * self.T * self.T + self.T)

    def status_word(self):
        """Device-side error bits of the last qagnn_graph_prep on this buffer (0 = all indices in range)."""
        return self.array("status")[0:1]

    def array(self, name):
        """int32 view of one prep array (tests / inspection)."""
        Ep = self.N + self.E
        n = {"rowptr_src": self.N + 1, "rowptr_tgt": self.N + 1, "status": 4}.get(name, Ep)
        off = getattr(self.layout, name)
        return self.buf[off:off + 4 * n].view(torch.int32)

    def edge_index_prime(self):
        """edge_index' = [edge_index | self loops] as int64 [2, E+N] (modeling_qagnn.py:436-438)."""
        return torch.stack([self.array("src"), self.array("tgt")]).long()


def _edge_encoder_params(enc):
    lin0, bn, lin3 = enc[0], enc[1], enc[3]
    ts = [lin0.weight, lin0.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, lin3.weight, lin3.bias]
    return ts


def _layer_params(layer):
    bn = layer.mlp[1]
    return [layer.linear_key.weight, layer.linear_key.bias, layer.linear_msg.weight, layer.linear_msg.bias,
            layer.linear_query.weight, layer.linear_query.bias, layer.mlp[0].weight, layer.mlp[0].bias,
            bn.weight, bn.bias, bn.running_mean, bn.running_var, layer.mlp[3].weight, layer.mlp[3].bias]


class _FoldedWeights:
    """Folded-parameter blob (qagnn_fold_weights), refreshed whenever a source tensor changes."""

    def __init__(self):
        self.key = None
        self.blob = None
        self._keep = None

    def invalidate(self):
        """Forget the folded blob.  Needed after parameter edits that bypass autograd's version counter
        (`p.data.normal_()`, `bn.running_mean.data.fill_()`): the cache key is (data_ptr, _version, shape)."""
        self.key = None

    def get(self, shape, enc, layers, mp_tensors, device):
        lib = _lib.load()
        srcs = _edge_encoder_params(enc)
        for l in layers:
            srcs += _layer_params(l)
        if mp_tensors is not None:
            srcs += mp_tensors
        key = (_version_key(srcs), shape.D, shape.H, shape.T, shape.R, shape.k, str(device))
        if key == self.key:
            return self.blob
        cont = [_lib.f32c(t, "parameter") for t in srcs]
        it = iter(cont)
        ee = _lib.EdgeEncoderParams(*[t.data_ptr() for t in [next(it) for _ in range(8)]])
        larr = (_lib.LayerParams * max(len(layers), 1))()
        for i in range(len(layers)):
            larr[i] = _lib.LayerParams(*[t.data_ptr() for t in [next(it) for _ in range(14)]])
        mp = None
        if mp_tensors is not None:
            mp = _lib.MPParams(*[t.data_ptr() for t in [next(it) for _ in range(9)]])
        nbytes = lib.qagnn_fold_bytes(C.byref(shape))
        if nbytes == 0:
            raise _lib.QagnnError("qagnn_fold_bytes: invalid shape")
        blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            st = lib.qagnn_fold_weights(C.byref(shape), C.byref(ee), larr, C.byref(mp) if mp is not None else None,
                                        _lib.ptr(blob), nbytes, _lib.stream_ptr(device))
        _lib.check(st, "qagnn_fold_weights")
        self.key, self.blob, self._keep = key, blob, cont
        return blob


class GATConvE(nn.Module):
    """
    Args:
        emb_dim (int): dimensionality of GNN hidden states
        n_ntype (int): number of node types (e.g. 4)
        n_etype (int): number of edge relation types (e.g. 38)
    Same constructor as the reference (modeling_qagnn.py:387): `edge_encoder` is the shared
    Sequential(Linear, BatchNorm1d, ReLU, Linear); `head_count` defaults to 4.
    """

    def __init__(self, args, emb_dim, n_ntype, n_etype, edge_encoder, head_count=4, aggr="add"):
        super().__init__()
        if aggr != "add":
            raise ValueError("GATConvE only supports aggr='add' (the only mode the reference uses)")
        if emb_dim % 2 != 0 or emb_dim % head_count != 0:
            raise ValueError("emb_dim must be even and divisible by head_count")
        self.args = args
        self.emb_dim = emb_dim
        self.n_ntype, self.n_etype = n_ntype, n_etype
        self.edge_encoder = edge_encoder
        self.head_count = head_count
        self.dim_per_head = emb_dim // head_count
        self.linear_key = nn.Linear(3 * emb_dim, head_count * self.dim_per_head)
        self.linear_msg = nn.Linear(3 * emb_dim, head_count * self.dim_per_head)
        self.linear_query = nn.Linear(2 * emb_dim, head_count * self.dim_per_head)
        self._alpha = None
        self.mlp = nn.Sequential(nn.Linear(emb_dim, emb_dim), nn.BatchNorm1d(emb_dim), nn.ReLU(),
                                 nn.Linear(emb_dim, emb_dim))
        self.check_indices = True
        self._folded = _FoldedWeights()
        self._ws = _DeviceBlob()
        self._prep_cache = (None, None)

    def _prep(self, edge_index, edge_type, node_type):
        key = _version_key([edge_index, edge_type, node_type])
        if self._prep_cache[0] != key:
            prep = GraphPrep(edge_index, edge_type, node_type, self.n_ntype, self.n_etype, 0, self.check_indices)
            self._prep_cache = (key, prep)
        return self._prep_cache[1]

    def forward(self, x, edge_index, edge_type, node_type, node_feature_extra, return_attention_weights=False,
                prep=None, return_aggr=False):
        """x [N, emb_dim]; edge_index [2, E]; edge_type [E]; node_type [N]; node_feature_extra [N, emb_dim].
        Returns out [N, emb_dim], or (out, (edge_index' [2,E+N], alpha [E+N, heads])) with
        return_attention_weights=True (alpha = softmax before the out-degree rescale, :473)."""
        if self.training:  # dropout-free layer, BatchNorm batch statistics, autograd through the CUDA message passing
            from . import training as TR
            if prep is None:
                prep = self._prep(edge_index, edge_type, node_type)
            tab = TR.edge_table_train(self.edge_encoder, TR.combo_onehot_table(self.n_ntype, self.n_etype, x.device),
                                      prep.combo_count(), 1)
            return TR.gatconve_train(self, x, node_feature_extra, node_type, prep, tab, return_attention_weights)
        _warn_if_grad_wanted("GATConvE", x, node_feature_extra)
        lib = _lib.load()
        xc = _lib.f32c(x, "x")
        ex = _lib.f32c(node_feature_extra, "node_feature_extra")
        dev = xc.device
        N, D = xc.shape
        if D != self.emb_dim or ex.shape != xc.shape:
            raise ValueError("x / node_feature_extra must be [N, emb_dim]")
        if prep is None:
            prep = self._prep(edge_index, edge_type, node_type)
        if prep.N != N:
            raise ValueError("node_type and x disagree on the number of nodes")
        shape = _lib.Shape(N, prep.E, D, self.head_count, self.n_ntype, self.n_etype, 1, prep.n_per_graph)
        folded = self._folded.get(shape, self.edge_encoder, [self], None, dev)
        ws_bytes = lib.qagnn_forward_workspace_bytes(C.byref(shape))
        ws = self._ws.get(ws_bytes, dev)
        out = torch.empty_like(xc)
        alpha = torch.empty(N + prep.E, self.head_count, dtype=torch.float32, device=dev) if return_attention_weights else None
        aggr = torch.empty_like(xc) if return_aggr else None
        with torch.cuda.device(dev):
            st = lib.qagnn_gatconve_forward(C.byref(shape), 0, _lib.ptr(xc), _lib.ptr(ex), _lib.ptr(prep.buf),
                                            _lib.ptr(folded), _lib.ptr(out), _lib.ptr(alpha), _lib.ptr(aggr),
                                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(st, "qagnn_gatconve_forward")
        self._alpha = None
        res = out
        if return_attention_weights:
            res = (out, (prep.edge_index_prime(), alpha))
        if return_aggr:
            return res, aggr
        return res


class QAGNN_Message_Passing(nn.Module):
    def __init__(self, args, k, n_ntype, n_etype, input_size, hidden_size, output_size, dropout=0.1):
        super().__init__()
        if input_size != output_size or input_size != hidden_size:
            raise ValueError("QAGNN_Message_Passing requires input_size == hidden_size == output_size")
        self.args = args
        self.n_ntype, self.n_etype = n_ntype, n_etype
        self.hidden_size = hidden_size
        self.emb_node_type = nn.Linear(self.n_ntype, hidden_size // 2)
        self.basis_f = "sin"
        self.emb_score = nn.Linear(hidden_size // 2, hidden_size // 2)
        self.edge_encoder = nn.Sequential(nn.Linear(n_etype + 1 + n_ntype * 2, hidden_size), nn.BatchNorm1d(hidden_size),
                                          nn.ReLU(), nn.Linear(hidden_size, hidden_size))
        self.k = k
        self.gnn_layers = nn.ModuleList([GATConvE(args, hidden_size, n_ntype, n_etype, self.edge_encoder) for _ in range(k)])
        self.Vh = nn.Linear(input_size, output_size)
        self.Vx = nn.Linear(hidden_size, output_size)
        self.activation = GELU()
        self.dropout = nn.Dropout(dropout)
        self.dropout_rate = dropout
        # 1.1^j in float32, computed on the host exactly as torch.pow(1.1, arange) does in the
        # reference (:70-71); uploaded as a constant so host and device agree bit for bit
        self.register_buffer("_score_basis", torch.pow(1.1, torch.arange(hidden_size // 2).float()), persistent=False)
        self.check_indices = True
        # use_cuda_graph: capture graph prep + the whole k-layer launch sequence (~60 kernels) into one CUDA graph per
        # (input buffers, weights) and replay it; the returned tensor is then a static buffer that the next replay
        # overwrites.  Index validation needs a stream sync, so it is skipped inside the captured region.
        self.use_cuda_graph = False
        self._graphs = {}
        self._folded = _FoldedWeights()
        # a captured graph bakes in the workspace / folded-blob addresses: drop the graphs before either is replaced
        self._ws = _DeviceBlob(on_replace=self._graphs.clear)
        self._deferred_status = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self):
        """Drops the folded weights and every captured CUDA graph.  Called automatically by load_state_dict() and
        .to()/.cuda(); call it yourself after in-place edits through `.data` (e.g. `p.data.normal_()`, EMA swaps),
        which do not bump the tensor version the caches are keyed on."""
        self._folded.invalidate()
        self._graphs.clear()

    def _apply(self, fn, *a, **kw):
        self.invalidate()
        return super()._apply(fn, *a, **kw)

    # -- helpers ------------------------------------------------------------------------------
    def _shape(self, N, E, n_per_graph):
        return _lib.Shape(N, E, self.hidden_size, self.gnn_layers[0].head_count if self.k else 4, self.n_ntype,
                          self.n_etype, self.k, n_per_graph)

    def _mp_tensors(self):
        return [self.emb_node_type.weight, self.emb_node_type.bias, self.emb_score.weight, self.emb_score.bias,
                self.Vh.weight, self.Vh.bias, self.Vx.weight, self.Vx.bias, self._score_basis]

    def _raise_deferred_status(self):
        """CUDA-graph replays cannot synchronise to validate indices: every replay ORs graph prep's status word into a
        device accumulator, which is copied to pinned memory asynchronously and inspected here once that copy has landed."""
        if self._deferred_status is None:
            return
        host, ev, accum = self._deferred_status
        if not ev.query():  # never block the host on the GPU here: the OR-accumulated word is looked at again later
            return
        self._deferred_status = None
        if int(host.item()) != 0:
            accum.zero_()
            raise IndexError("qagnn_b200: the previous CUDA-graph replay saw edge_index / edge_type / node_type values out "
                             f"of range or crossing a sub-graph boundary (status bits {int(host.item())}); its output was "
                             "computed from clamped indices")

    def check_batch(self, A, node_type):
        """Validates the indices of a batch once (synchronising); raises IndexError like forward() with check_indices."""
        n = node_type.size(1) if node_type.dim() == 2 else 0
        GraphPrep(A[0], A[1], node_type.reshape(-1), self.n_ntype, self.n_etype, n, True, allow_general_fallback=True,
                  graph_ptr=getattr(A, "graph_ptr_dev", None), max_edges=getattr(A, "max_edges", -1))

    def prepare_graph(self, edge_index, edge_type, node_type):
        """Builds the layer-invariant graph workspace; pass it back via forward(..., prep=) to amortise
        it over repeated forwards on the same batch."""
        n_per_graph = node_type.size(1) if node_type.dim() == 2 else 0
        return GraphPrep(edge_index, edge_type, node_type, self.n_ntype, self.n_etype, n_per_graph, self.check_indices)

    def mp_helper(self, _X, edge_index, edge_type, _node_type, _node_feature_extra):
        for layer in self.gnn_layers:
            _X = self.activation(layer(_X, edge_index, edge_type, _node_type, _node_feature_extra))
        return _X

    def node_feature_extra(self, node_type, node_score):
        """[B*n, D] = [GELU(emb_node_type(onehot(type))) ‖ GELU(emb_score(sin(1.1^j * score)))] (:62-73,86)."""
        lib = _lib.load()
        nt = _lib.i64c(node_type.reshape(-1), "node_type")
        sc = _lib.f32c(node_score.reshape(-1), "node_score")
        dev = nt.device
        shape = self._shape(nt.numel(), 0, 0)
        folded = self._folded.get(shape, self.edge_encoder, list(self.gnn_layers), self._mp_tensors(), dev)
        ws = self._ws.get(lib.qagnn_forward_workspace_bytes(C.byref(shape)), dev)
        out = torch.empty(nt.numel(), self.hidden_size, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = lib.qagnn_node_feature_extra(C.byref(shape), _lib.ptr(nt), _lib.ptr(sc), _lib.ptr(folded), _lib.ptr(out),
                                              _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(st, "qagnn_node_feature_extra")
        return out

    def forward(self, H, A, node_type, node_score, cache_output=False, prep=None, return_layers=False):
        """
        H: tensor of shape (batch_size, n_node, d_node) — node features from the previous layer
        A: (edge_index [2, total_E], edge_type [total_E]) of the batched graph
        node_type: long tensor (batch_size, n_node): 0 question entity, 1 answer entity, 2 other, 3 context
        node_score: tensor (batch_size, n_node, 1)
        returns (batch_size, n_node, d_node)
        """
        self._raise_deferred_status()
        if self.training:  # dropout, BatchNorm batch statistics, autograd through the CUDA message passing (training.py)
            from . import training as TR
            if prep is None:
                prep = GraphPrep(A[0], A[1], node_type.reshape(-1), self.n_ntype, self.n_etype, 0, self.check_indices)
            return TR.mp_forward_train(self, H, A, node_type, node_score, prep)
        _warn_if_grad_wanted("QAGNN_Message_Passing", H, node_score)
        if self.use_cuda_graph and prep is None and not return_layers:
            return self._forward_graphed(H, A, node_type, node_score)
        lib = _lib.load()
        Hc = _lib.f32c(H, "H")
        if Hc.dim() != 3 or Hc.size(2) != self.hidden_size:
            raise ValueError("H must be [batch, n_node, hidden_size]")
        B, n, D = Hc.shape
        dev = Hc.device
        nt = _lib.i64c(node_type.reshape(-1), "node_type")
        sc = _lib.f32c(node_score.reshape(-1), "node_score")
        if nt.numel() != B * n or sc.numel() != B * n:
            raise ValueError("node_type / node_score must be [batch, n_node(, 1)]")
        edge_index, edge_type = A  # a (edge_index, edge_type) pair, or a qagnn_b200.data.PackedAdj (iterates as that pair)
        if prep is None:
            gp, me = getattr(A, "graph_ptr_dev", None), getattr(A, "max_edges", -1)
            if gp is not None and getattr(A, "n_nodes", n) != n:
                gp = None  # packed for another node count: its graph_ptr still holds, its offsets do not -> general prep
            prep = GraphPrep(edge_index, edge_type, nt, self.n_ntype, self.n_etype, n, self.check_indices,
                             allow_general_fallback=True, graph_ptr=gp, max_edges=me)
        self._last_prep = prep
        if prep.N != B * n:
            raise ValueError("graph workspace was built for a different number of nodes")
        shape = self._shape(B * n, prep.E, prep.n_per_graph)
        folded = self._folded.get(shape, self.edge_encoder, list(self.gnn_layers), self._mp_tensors(), dev)
        ws = self._ws.get(lib.qagnn_forward_workspace_bytes(C.byref(shape)), dev)
        out = torch.empty_like(Hc)
        layers = torch.empty(self.k, B * n, D, dtype=torch.float32, device=dev) if return_layers else None
        with torch.cuda.device(dev):
            st = lib.qagnn_mp_forward(C.byref(shape), _lib.ptr(Hc), _lib.ptr(nt), _lib.ptr(sc), _lib.ptr(prep.buf),
                                      _lib.ptr(folded), _lib.ptr(out), _lib.ptr(layers), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_ptr(dev))
        _lib.check(st, "qagnn_mp_forward")
        if return_layers:
            return out, layers
        return out


def _mp_forward_graphed(self, H, A, node_type, node_score):
    """CUDA-graph replay of forward() for inputs that live in the same device buffers as when it was captured."""
    tensors = [H, A[0], A[1], node_type, node_score]
    if getattr(A, "graph_ptr_dev", None) is not None:
        tensors.append(A.graph_ptr_dev)
    key = (tuple((t.data_ptr(), tuple(t.shape), t.dtype) for t in tensors),
           _version_key(list(self.parameters()) + list(self.buffers())))
    entry = self._graphs.get(key)
    if entry is None:
        saved, self.use_cuda_graph = self.use_cuda_graph, False
        check, self.check_indices = self.check_indices, False
        try:
            if check:  # validate once, eagerly, before trusting the capture
                self.check_batch(A, node_type)
            side = torch.cuda.Stream(device=H.device)
            side.wait_stream(torch.cuda.current_stream(H.device))
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up: folds the weights, sizes the workspaces, sets kernel attributes
                    self.forward(H, A, node_type, node_score)
            torch.cuda.current_stream(H.device).wait_stream(side)
            accum = torch.zeros(1, dtype=torch.int32, device=H.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward(H, A, node_type, node_score)
                accum.bitwise_or_(self._last_prep.status_word())
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            # the entry keeps alive everything whose address the graph baked in: inputs, workspace, folded blob, prep
            entry = self._graphs[key] = (graph, out, tensors, self._ws.buf, self._folded.blob, self._last_prep, accum,
                                         torch.zeros(1, dtype=torch.int32).pin_memory())
        finally:
            self.use_cuda_graph, self.check_indices = saved, check
    entry[0].replay()
    if self.check_indices:  # later batches written into the same buffers: checked at the next call, without a sync here
        entry[7].copy_(entry[6], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(H.device))
        self._deferred_status = (entry[7], ev, entry[6])
    return entry[1]


QAGNN_Message_Passing._forward_graphed = _mp_forward_graphed


class QAGNN(nn.Module):
    """Decoder around the message passing (modeling_qagnn.py:99-189); plain PyTorch on the caller side."""

    def __init__(self, args, k, n_ntype, n_etype, sent_dim, n_concept, concept_dim, concept_in_dim, n_attention_head,
                 fc_dim, n_fc_layer, p_emb, p_gnn, p_fc, pretrained_concept_emb=None, freeze_ent_emb=True,
                 init_range=0.02):
        super().__init__()
        self.init_range = init_range
        self.concept_emb = CustomizedEmbedding(concept_num=n_concept, concept_out_dim=concept_dim, use_contextualized=False,
                                               concept_in_dim=concept_in_dim, pretrained_concept_emb=pretrained_concept_emb,
                                               freeze_ent_emb=freeze_ent_emb)
        self.svec2nvec = nn.Linear(sent_dim, concept_dim)
        self.concept_dim = concept_dim
        self.activation = GELU()
        self.gnn = QAGNN_Message_Passing(args, k=k, n_ntype=n_ntype, n_etype=n_etype, input_size=concept_dim,
                                         hidden_size=concept_dim, output_size=concept_dim, dropout=p_gnn)
        self.pooler = MultiheadAttPoolLayer(n_attention_head, sent_dim, concept_dim)
        self.fc = MLP(concept_dim + sent_dim + concept_dim, fc_dim, 1, n_fc_layer, p_fc, layer_norm=True)
        self.dropout_e = nn.Dropout(p_emb)
        self.dropout_fc = nn.Dropout(p_fc)
        if init_range > 0:
            self.apply(self._init_weights)

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.init_range)
            if getattr(module, "bias", None) is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    def pooled_features(self, sent_vecs, concept_ids, node_type_ids, node_scores, adj_lengths, adj, emb_data=None,
                        cache_output=False):
        """Everything up to (and excluding) the answer MLP: returns (concat [B, 2*concept_dim+sent_dim], pool_attn).
        Split out so a data-parallel caller can all-gather `concat` before `fc` (qagnn_b200.distributed)."""
        n = node_scores.size(1)
        pos = torch.arange(n, device=node_scores.device)
        fused_head = (not self.training and not torch.is_grad_enabled() and emb_data is None and node_scores.is_cuda
                      and not self.concept_emb.use_contextualized and self.concept_dim % 4 == 0
                      and concept_ids.device == node_scores.device and sent_vecs.dtype == torch.float32)
        if fused_head:
            # one kernel: context-node injection + concept rows gathered from the pre-projected table + score
            # normalisation (:153-167); the [B, sent_dim] svec2nvec GEMM stays a library call
            lib = _lib.load()
            B_ = concept_ids.size(0)
            table = self.concept_emb.projected_table()
            ctx = self.activation(self.svec2nvec(sent_vecs)).contiguous()
            cid = _lib.i64c(concept_ids, "concept_ids")
            sc = _lib.f32c(node_scores.reshape(B_, n), "node_scores")
            al = _lib.i64c(adj_lengths, "adj_lengths")
            gnn_input = torch.empty(B_, n, self.concept_dim, dtype=torch.float32, device=sc.device)
            s = torch.empty(B_, n, dtype=torch.float32, device=sc.device)
            with torch.cuda.device(sc.device):
                st = lib.qagnn_decoder_head(B_, n, self.concept_dim, _lib.ptr(cid), table.size(0), _lib.ptr(table), _lib.ptr(ctx),
                                            _lib.ptr(sc), _lib.ptr(al), _lib.ptr(gnn_input), _lib.ptr(s), _lib.stream_ptr(sc.device))
            _lib.check(st, "qagnn_decoder_head")
            s = s.unsqueeze(2)
        else:
            gnn_input0 = self.activation(self.svec2nvec(sent_vecs)).unsqueeze(1)
            gnn_input1 = self.concept_emb(concept_ids[:, 1:] - 1, emb_data).to(node_type_ids.device)
            gnn_input = self.dropout_e(torch.cat([gnn_input0, gnn_input1], dim=1))
            valid = (pos < adj_lengths.unsqueeze(1)).float()
            s = -node_scores
            s = (s - s[:, 0:1, :]).squeeze(2) * valid
            mean_norm = s.abs().sum(dim=1) / adj_lengths
            s = (s / (mean_norm.unsqueeze(1) + 1e-05)).unsqueeze(2)

        gnn_output = self.gnn(gnn_input, adj, node_type_ids, s)
        fused = self.pooler.pool_concat(sent_vecs, gnn_output, node_type_ids, adj_lengths) if not self.training else None
        if fused is not None:  # pool mask + pooling + cat(graph_vecs, sent_vecs, Z) in one kernel (dropout_fc = identity in eval)
            if cache_output:
                self.concept_ids, self.adj, self.pool_attn = concept_ids, adj, fused[1]
            return fused
        Z_vecs = gnn_output[:, 0]
        mask = (pos >= adj_lengths.unsqueeze(1)) | (node_type_ids == 3)
        mask[:, 0] = mask[:, 0] & ~mask.all(1)  # a fully masked row keeps node 0 (:176; written without a host sync)
        graph_vecs, pool_attn = self.pooler(sent_vecs, gnn_output, mask)
        if cache_output:
            self.concept_ids, self.adj, self.pool_attn = concept_ids, adj, pool_attn
        concat = self.dropout_fc(torch.cat((graph_vecs, sent_vecs, Z_vecs), 1))
        return concat, pool_attn

    def forward(self, sent_vecs, concept_ids, node_type_ids, node_scores, adj_lengths, adj, emb_data=None,
                cache_output=False):
        """sent_vecs (B, dim_sent); concept_ids (B, n_node); adj = (edge_index, edge_type);
        adj_lengths (B,); node_type_ids (B, n_node); node_scores (B, n_node, 1).  Returns (logits (B,1), pool_attn)."""
        concat, pool_attn = self.pooled_features(sent_vecs, concept_ids, node_type_ids, node_scores, adj_lengths, adj,
                                                 emb_data, cache_output)
        return self.fc(concat), pool_attn


class LM_QAGNN(nn.Module):
    """Text encoder + QAGNN decoder with the reference's positional-input forward (modeling_qagnn.py:192-251)."""

    def __init__(self, args, model_name, k, n_ntype, n_etype, n_concept, concept_dim, concept_in_dim, n_attention_head,
                 fc_dim, n_fc_layer, p_emb, p_gnn, p_fc, pretrained_concept_emb=None, freeze_ent_emb=True,
                 init_range=0.0, encoder_config={}):
        super().__init__()
        from .modeling_encoder import TextEncoder
        self.encoder = TextEncoder(model_name, **encoder_config)
        self.decoder = QAGNN(args, k, n_ntype, n_etype, self.encoder.sent_dim, n_concept, concept_dim, concept_in_dim,
                             n_attention_head, fc_dim, n_fc_layer, p_emb, p_gnn, p_fc,
                             pretrained_concept_emb=pretrained_concept_emb, freeze_ent_emb=freeze_ent_emb,
                             init_range=init_range)

    def forward(self, *inputs, layer_id=-1, cache_output=False, detail=False):
        """inputs = (*lm_inputs, concept_ids, node_type_ids, node_scores, adj_lengths, edge_index, edge_type) with a
        leading (batch, num_choice) pair of dims on the tensors and [batch][num_choice] nested lists for the last two
        (or a PackedAdj, see qagnn_b200.data).  Returns (logits [batch, num_choice], pool_attn)."""
        from .data import PackedAdj
        bs, nc = inputs[0].size(0), inputs[0].size(1)
        edge_index_orig, edge_type_orig = inputs[-2:]
        flat = [x.reshape(x.size(0) * x.size(1), *x.size()[2:]) for x in inputs[:-2]]
        *lm_inputs, concept_ids, node_type_ids, node_scores, adj_lengths = flat
        if isinstance(edge_index_orig, PackedAdj):  # pre-packed batch (qagnn_b200.data.pack_adj); edge_type slot unused
            adj = edge_index_orig.to(node_type_ids.device)  # keeps graph_ptr: graph prep is then one launch per batch
        else:
            edge_index, edge_type = self.batch_graph(sum(edge_index_orig, []), sum(edge_type_orig, []), concept_ids.size(1))
            adj = (edge_index.to(node_type_ids.device), edge_type.to(node_type_ids.device))
        sent_vecs, all_hidden_states = self.encoder(*lm_inputs, layer_id=layer_id)
        logits, attn = self.decoder(sent_vecs.to(node_type_ids.device), concept_ids, node_type_ids, node_scores,
                                    adj_lengths, adj, emb_data=None, cache_output=cache_output)
        logits = logits.view(bs, nc)
        if not detail:
            return logits, attn
        return (logits, attn, concept_ids.view(bs, nc, -1), node_type_ids.view(bs, nc, -1), edge_index_orig,
                edge_type_orig)

    def batch_graph(self, edge_index_init, edge_type_init, n_nodes):
        """list of [2,E_i] / [E_i] -> one [2, total_E] / [total_E] with node ids offset by i*n_nodes (:244-251)."""
        n_examples = len(edge_index_init)
        offsets = torch.repeat_interleave(
            torch.arange(n_examples, device=edge_index_init[0].device) * n_nodes,
            torch.tensor([e.size(1) for e in edge_index_init], device=edge_index_init[0].device))
        edge_index = torch.cat(edge_index_init, dim=1) + offsets.unsqueeze(0)
        edge_type = torch.cat(edge_type_init, dim=0)
        return edge_index, edge_type
