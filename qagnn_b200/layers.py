"""Small host-side layers around the hot path (the caller side of SURVEY.md §8b), kept as plain
PyTorch.  Constructor arguments and state_dict key names follow the reference's utils/layers.py
so checkpoints load unchanged:
    GELU                    utils/layers.py:10-22   (tanh approximation)
    MLP                     utils/layers.py:47-87   ('layers.{i}-Linear', '{i}-LayerNorm', ...)
    MultiheadAttPoolLayer   utils/layers.py:324-371 (+ MatrixVectorScaledDotProductAttention :276-299)
    CustomizedEmbedding     utils/layers.py:571-607
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def gelu(x):
    """tanh-approximation GELU, the form the reference uses everywhere (utils/layers.py:10-14)."""
    return F.gelu(x, approximate="tanh")


class GELU(nn.Module):
    def forward(self, x):
        return gelu(x)


class MLP(nn.Module):
    """num_layers hidden blocks (Linear, Dropout, [BatchNorm1d | LayerNorm], activation) + output Linear."""
    _ACT = {"gelu": GELU, "relu": nn.ReLU, "tanh": nn.Tanh}

    def __init__(self, input_size, hidden_size, output_size, num_layers, dropout, batch_norm=False,
                 init_last_layer_bias_to_zero=False, layer_norm=False, activation="gelu"):
        super().__init__()
        if batch_norm and layer_norm:
            raise ValueError("batch_norm and layer_norm are mutually exclusive")
        self.input_size, self.hidden_size, self.output_size = input_size, hidden_size, output_size
        self.num_layers, self.dropout, self.batch_norm, self.layer_norm = num_layers, dropout, batch_norm, layer_norm
        blocks = nn.Sequential()
        width = input_size
        for i in range(num_layers):
            blocks.add_module(f"{i}-Linear", nn.Linear(width, hidden_size))
            blocks.add_module(f"{i}-Dropout", nn.Dropout(dropout))
            if batch_norm:
                blocks.add_module(f"{i}-BatchNorm1d", nn.BatchNorm1d(hidden_size))
            if layer_norm:
                blocks.add_module(f"{i}-LayerNorm", nn.LayerNorm(hidden_size))
            blocks.add_module(f"{i}-{activation}", self._ACT[activation.lower()]())
            width = hidden_size
        blocks.add_module(f"{num_layers}-Linear", nn.Linear(width, output_size))
        self.layers = blocks
        if init_last_layer_bias_to_zero:
            self.layers[-1].bias.data.zero_()

    def forward(self, x):
        return self.layers(x)


class MultiheadAttPoolLayer(nn.Module):
    """Pools k [b, l, d_k_original] with a query q [b, d_q_original]: n_head scaled dot-product
    attentions over the l positions, masked positions excluded.  Returns (pooled [b, n_head*d_v],
    attn [n_head*b, l]) with the reference's head-major attn layout."""

    def __init__(self, n_head, d_q_original, d_k_original, dropout=0.1):
        super().__init__()
        if d_k_original % n_head != 0:
            raise ValueError("d_k_original must be divisible by n_head")
        self.n_head = n_head
        self.d_k = self.d_v = d_k_original // n_head
        self.w_qs = nn.Linear(d_q_original, n_head * self.d_k)
        self.w_ks = nn.Linear(d_k_original, n_head * self.d_k)
        self.w_vs = nn.Linear(d_k_original, n_head * self.d_v)
        nn.init.normal_(self.w_qs.weight, mean=0, std=math.sqrt(2.0 / (d_q_original + self.d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=math.sqrt(2.0 / (d_k_original + self.d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=math.sqrt(2.0 / (d_k_original + self.d_v)))
        self.temperature = math.sqrt(self.d_k)
        self.attn_dropout = nn.Dropout(0.1)  # MatrixVectorScaledDotProductAttention's own dropout
        self.dropout = nn.Dropout(dropout)
        self._fused_ok = True

    def pool_concat(self, q, k, node_type, adj_lengths):
        """Eval-mode decoder tail in ONE kernel (qagnn_decoder_tail): derives the pool mask from adj_lengths / node_type,
        pools, and returns (cat(pooled, q, k[:, 0]) [b, 2*d_k_original + d_q_original], attn) — modeling_qagnn.py:172-187.
        Returns None when the fused kernel cannot take the shape (caller falls back to mask + forward + cat)."""
        if (self.training or not k.is_cuda or k.dtype != torch.float32 or not self._fused_ok
                or (torch.is_grad_enabled() and (q.requires_grad or k.requires_grad))):
            return None
        from . import _lib
        lib = _lib.load()
        b, l, D = k.shape
        S = q.size(1)
        qc = _lib.f32c(q, "q")
        qs = self.w_qs(qc).detach().contiguous()
        kc = _lib.f32c(k, "k")
        concat = torch.empty(b, 2 * D + S, dtype=torch.float32, device=k.device)
        attn = torch.empty(self.n_head * b, l, dtype=torch.float32, device=k.device)
        with torch.cuda.device(k.device):
            st = lib.qagnn_decoder_tail(b, l, D, self.n_head, S, _lib.ptr(kc), _lib.ptr(qs), _lib.ptr(_lib.i64c(node_type, "node_type")),
                                        _lib.ptr(_lib.i64c(adj_lengths, "adj_lengths")), _lib.ptr(qc),
                                        _lib.ptr(self.w_ks.weight.detach()), _lib.ptr(self.w_ks.bias.detach()),
                                        _lib.ptr(self.w_vs.weight.detach()), _lib.ptr(self.w_vs.bias.detach()),
                                        _lib.ptr(concat), _lib.ptr(attn), _lib.stream_ptr(k.device))
        if st == -5:
            self._fused_ok = False
            return None
        _lib.check(st, "qagnn_decoder_tail")
        return concat, attn

    def forward(self, q, k, mask=None):
        b, l, _ = k.shape
        nh, dk, dv = self.n_head, self.d_k, self.d_v
        wants_grad = torch.is_grad_enabled() and (q.requires_grad or k.requires_grad)  # the fused kernel is outside autograd
        if (k.is_cuda and not self.training and not wants_grad and mask is not None and k.dtype == torch.float32
                and self.w_ks.weight.shape[0] == k.shape[2] and self._fused_ok):
            # fused kernel (one read of k, no ks/vs GEMMs): qagnn_attention_pool
            from . import _lib
            lib = _lib.load()
            qs = self.w_qs(q).contiguous()
            kc = _lib.f32c(k, "k")
            m8 = mask.to(torch.uint8).contiguous()
            pooled = torch.empty(b, nh * dv, dtype=torch.float32, device=k.device)
            attn = torch.empty(nh * b, l, dtype=torch.float32, device=k.device)
            with torch.cuda.device(k.device):
                st = lib.qagnn_attention_pool(b, l, k.shape[2], nh, _lib.ptr(kc), _lib.ptr(qs.detach()), _lib.ptr(m8),
                                              _lib.ptr(self.w_ks.weight.detach()), _lib.ptr(self.w_ks.bias.detach()),
                                              _lib.ptr(self.w_vs.weight.detach()), _lib.ptr(self.w_vs.bias.detach()),
                                              _lib.ptr(pooled), _lib.ptr(attn), _lib.stream_ptr(k.device))
            if st == -5:  # QAGNN_ERR_UNSUPPORTED (node tile does not fit in shared memory): the einsum path below
                self._fused_ok = False
            else:
                _lib.check(st, "qagnn_attention_pool")
                return pooled, attn
        qs = self.w_qs(q).view(b, nh, dk)
        ks = self.w_ks(k).view(b, l, nh, dk)
        vs = self.w_vs(k).view(b, l, nh, dv)
        logits = torch.einsum("bhd,blhd->hbl", qs, ks) / self.temperature
        if mask is not None:
            logits = logits.masked_fill(mask.unsqueeze(0), float("-inf"))
        attn = self.attn_dropout(torch.softmax(logits, dim=2))
        pooled = torch.einsum("hbl,blhd->bhd", attn, vs).reshape(b, nh * dv)
        return self.dropout(pooled), attn.reshape(nh * b, l)


class CustomizedEmbedding(nn.Module):
    """Concept embedding table with an optional Linear+GELU projection when the table width differs
    from the GNN width."""

    def __init__(self, concept_num, concept_in_dim, concept_out_dim, use_contextualized=False,
                 pretrained_concept_emb=None, freeze_ent_emb=True, scale=1.0, init_range=0.02):
        super().__init__()
        self.scale = scale
        self.use_contextualized = use_contextualized
        if not use_contextualized:
            self.emb = nn.Embedding(concept_num, concept_in_dim)
            if pretrained_concept_emb is not None:
                self.emb.weight.data.copy_(pretrained_concept_emb)
            else:
                self.emb.weight.data.normal_(mean=0.0, std=init_range)
            if freeze_ent_emb:
                for p in self.emb.parameters():
                    p.requires_grad = False
        if concept_in_dim != concept_out_dim:
            self.cpt_transform = nn.Linear(concept_in_dim, concept_out_dim)
            self.activation = GELU()

    def _project(self, e):
        e = e * self.scale
        if not hasattr(self, "cpt_transform"):
            return e
        lin = self.cpt_transform
        if (e.is_cuda and not self.training and not torch.is_grad_enabled() and e.dtype == torch.float32
                and lin.in_features % 8 == 0 and lin.out_features >= 8):
            # GELU(cpt_transform(e)) on the tcgen05 split-bf16 GEMM with the activation fused in the epilogue: the
            # [B*(n-1), concept_in_dim] x [concept_in_dim, concept_dim] product is the largest dense op outside the GNN
            from . import ops
            try:
                out = ops.linear_bf16x3(e.reshape(-1, lin.in_features), lin.weight, lin.bias, act="gelu")
                return out.view(*e.shape[:-1], lin.out_features)
            except Exception as exc:  # tensor-core path not available for this shape/driver: plain PyTorch (same math)
                if "unsupported" not in str(exc).lower():
                    raise
        return self.activation(lin(e))

    def projected_table(self):
        """[concept_num, concept_out_dim] = the whole table pushed through scale / cpt_transform / GELU, cached per weight
        version (eval mode: the embedding is frozen and the projection fixed, so a forward becomes a pure row gather).
        `invalidate_table()` after in-place `.data` edits."""
        srcs = [self.emb.weight] + ([self.cpt_transform.weight, self.cpt_transform.bias] if hasattr(self, "cpt_transform") else [])
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in srcs)
        if getattr(self, "_ptab_key", None) != key:
            with torch.no_grad():
                w = self.emb.weight
                rows = [self._project(w[i:i + 65536]) for i in range(0, w.size(0), 65536)]  # chunks bound the GEMM workspace
                self._ptab = torch.cat(rows).contiguous()
            self._ptab_key = key
        return self._ptab

    def invalidate_table(self):
        self._ptab_key = None

    def forward(self, index, contextualized_emb=None):
        if contextualized_emb is not None:
            if index.size(0) != contextualized_emb.size(0):
                raise ValueError("batch size mismatch between index and contextualized_emb")
            table = self._project(contextualized_emb)
            return table.gather(1, index.unsqueeze(-1).expand(-1, -1, table.size(-1)))
        return self._project(self.emb(index))
