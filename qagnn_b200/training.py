"""Training-mode forward + backward of the QA-GNN message-passing path.

What the reference does in `qagnn.py:249-278` (`model.train()`, `loss.backward()`) for this path:
dropout after every layer (`modeling_qagnn.py:49,93`), BatchNorm batch statistics in the shared
`edge_encoder` (`:30`) and in every layer's `mlp` (`:408`), and autograd through
`GATConvE.message` / `propagate` (`:442,455-484`).

Split here:
  * the graph part — logits, per-source softmax, out-degree rescale, per-target sum and ITS GRADIENT — runs in the
    hand-written kernels of libqagnn_b200.so (`qagnn_mp_core_forward` / `qagnn_mp_core_backward`, csrc/message_passing.cu,
    csrc/mp_backward.cu) behind one `torch.autograd.Function`;
  * the dense linears, BatchNorm and dropout around it are ordinary PyTorch ops on the GPU (library GEMMs), so their
    gradients come from autograd; the same node-level factorisation as the eval path is used (K/M/Q projections per node,
    the edge encoder evaluated once per distinct one-hot "combo"), which is algebraically the reference's per-edge form.

BatchNorm of the edge encoder: the reference feeds it one row per edge (E+N rows, `:433`), but a row only depends on the
edge's combo, so the batch mean / variance are the combo-histogram-weighted moments of the C distinct rows; the running
statistics receive the k momentum updates the reference's k calls of the shared module apply.
"""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib


def combo_onehot_table(n_ntype, n_etype, device):
    """[C, R+1+2T] one-hot edge-feature rows, one per combo index used by graph prep
    (c = (etype*T + type_src)*T + type_tgt for real edges, R*T*T + type for the self loop; modeling_qagnn.py:419-432)."""
    T, R = n_ntype, n_etype
    Cn = R * T * T + T
    c = torch.arange(Cn, device=device)
    real = c < R * T * T
    et = torch.where(real, c // (T * T), torch.full_like(c, R))
    ts = torch.where(real, (c // T) % T, c - R * T * T)
    tt = torch.where(real, c % T, c - R * T * T)
    tab = torch.zeros(Cn, R + 1 + 2 * T, device=device)
    tab[c, et] = 1.0
    tab[c, R + 1 + ts] = 1.0
    tab[c, R + 1 + T + tt] = 1.0
    return tab


def edge_table_train(enc, onehot_tab, combo_count, k_calls):
    """tab[c] = edge_encoder(onehot(c)) with BATCH statistics over the E' edges (each combo weighted by its edge count),
    plus the running-statistics update of `k_calls` forward calls of the shared module."""
    lin0, bn, lin3 = enc[0], enc[1], enc[3]
    h = lin0(onehot_tab)                                        # [C, D]
    w = combo_count.to(h.dtype)
    n = w.sum()
    mean = (w[:, None] * h).sum(0) / n
    var = (w[:, None] * (h - mean) ** 2).sum(0) / n              # biased, as F.batch_norm normalises with
    y = (h - mean) * torch.rsqrt(var + bn.eps) * bn.weight + bn.bias
    with torch.no_grad():
        if bn.track_running_stats and bn.running_mean is not None:
            m = bn.momentum if bn.momentum is not None else 0.1
            keep = (1.0 - m) ** k_calls
            unbiased = var * (n / (n - 1.0)) if float(n) > 1 else var
            bn.running_mean.mul_(keep).add_(mean * (1.0 - keep))
            bn.running_var.mul_(keep).add_(unbiased * (1.0 - keep))
            bn.num_batches_tracked += k_calls
    return lin3(F.relu(y))


class _MPCore(torch.autograd.Function):
    """aggr = propagate(...) on node-level projections; forward and backward in libqagnn_b200.so."""

    # under torch.autocast (qagnn.py:91 runs the forward in fp16 autocast when --fp16 is set) the dense layers around this
    # function may hand it half tensors: the graph part always computes in fp32, like the reference's scatter ops do
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, qkm, ke, me, prep, shape_args, want_alpha):
        lib = _lib.load()
        qkm_c, ke_c, me_c = _lib.f32c(qkm, "qkm"), _lib.f32c(ke, "ke"), _lib.f32c(me, "me")
        N, E, D, H, T, R = shape_args
        dev = qkm_c.device
        shape = _lib.Shape(N, E, D, H, T, R, 1, 0)
        aggr = torch.empty(N, D, dtype=torch.float32, device=dev)
        alpha_s = torch.empty(N + E, H, dtype=torch.float32, device=dev)
        alpha = torch.empty(N + E, H, dtype=torch.float32, device=dev) if want_alpha else None
        scratch = torch.empty(N + E, H, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = lib.qagnn_mp_core_forward(C.byref(shape), _lib.ptr(prep.buf), _lib.ptr(qkm_c), _lib.ptr(ke_c), _lib.ptr(me_c),
                                           _lib.ptr(aggr), _lib.ptr(alpha_s), _lib.ptr(alpha), _lib.ptr(scratch),
                                           _lib.stream_ptr(dev))
        _lib.check(st, "qagnn_mp_core_forward")
        ctx.save_for_backward(qkm_c, ke_c, me_c, alpha_s)
        ctx.prep, ctx.shape_args = prep, shape_args
        if want_alpha:
            ctx.mark_non_differentiable(alpha)
            return aggr, alpha
        return aggr, None

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, d_aggr, _d_alpha):
        lib = _lib.load()
        qkm, ke, me, alpha_s = ctx.saved_tensors
        N, E, D, H, T, R = ctx.shape_args
        dev = qkm.device
        shape = _lib.Shape(N, E, D, H, T, R, 1, 0)
        g = _lib.f32c(d_aggr, "d_aggr")
        d_qkm = torch.empty_like(qkm)
        d_ke, d_me = torch.empty_like(ke), torch.empty_like(me)
        scratch = torch.empty(N + E, H, dtype=torch.float32, device=dev)
        order = ctx.prep.combo_order()
        with torch.cuda.device(dev):
            st = lib.qagnn_mp_core_backward(C.byref(shape), _lib.ptr(ctx.prep.buf), _lib.ptr(order), _lib.ptr(qkm), _lib.ptr(ke),
                                            _lib.ptr(me), _lib.ptr(alpha_s), _lib.ptr(g), _lib.ptr(d_qkm), _lib.ptr(d_ke),
                                            _lib.ptr(d_me), _lib.ptr(scratch), _lib.stream_ptr(dev))
        _lib.check(st, "qagnn_mp_core_backward")
        return d_qkm, d_ke, d_me, None, None, None


def gatconve_train(layer, x, extra, node_type, prep, tab, return_attention_weights=False):
    """GATConvE.forward in training mode (modeling_qagnn.py:411-452) given the edge table `tab` [C, D]."""
    D, H = layer.emb_dim, layer.head_count
    d = D // H
    x2 = torch.cat([x, extra], dim=1)                                        # :440
    wk, wm = layer.linear_key.weight, layer.linear_msg.weight
    q = layer.linear_query(x2) / (d ** 0.5)                                  # :466,:469
    kx = F.linear(x2, wk[:, :2 * D])                                         # node part of :464
    mx = F.linear(x2, wm[:, :2 * D])                                         # node part of :465
    ke = F.linear(tab, wk[:, 2 * D:], layer.linear_key.bias)                 # edge part of :464 (+ bias)
    me = F.linear(tab, wm[:, 2 * D:], layer.linear_msg.bias)                 # edge part of :465 (+ bias)
    qkm = torch.cat([q, kx, mx], dim=1)
    aggr, alpha = _MPCore.apply(qkm, ke, me, prep, (x.size(0), prep.E, D, H, layer.n_ntype, layer.n_etype),
                                return_attention_weights)
    out = layer.mlp(aggr)                                                    # :443 (BatchNorm batch statistics)
    if return_attention_weights:
        return out, (prep.edge_index_prime(), alpha)
    return out


def mp_forward_train(mod, H, A, node_type, node_score, prep):
    """QAGNN_Message_Passing.forward in training mode (modeling_qagnn.py:53-95)."""
    B, n, D = H.shape
    nt = node_type.reshape(-1)
    T_ = F.one_hot(nt, mod.n_ntype).to(H.dtype)
    type_emb = mod.activation(mod.emb_node_type(T_))                         # :65-66
    js = mod._score_basis.to(H.device)                                       # 1.1^j, float32 as the reference computes it
    Bs = torch.sin(js.view(1, -1) * node_score.reshape(-1, 1))               # :70-72
    score_emb = mod.activation(mod.emb_score(Bs))                            # :73
    extra = torch.cat([type_emb, score_emb], dim=1)                          # :86
    tab = edge_table_train(mod.edge_encoder, combo_onehot_table(mod.n_ntype, mod.n_etype, H.device), prep.combo_count(),
                           max(mod.k, 1))
    X = H.reshape(-1, D)
    for layer in mod.gnn_layers:                                             # mp_helper :45-50
        X = mod.activation(gatconve_train(layer, X, extra, nt, prep, tab))
        X = F.dropout(X, mod.dropout_rate, training=True)
    out = mod.activation(mod.Vh(H) + mod.Vx(X.view(B, n, D)))                # :92
    return mod.dropout(out)                                                  # :93
