#!/usr/bin/env python
"""Timings for the other BASELINE.json configs (cfg1 tiny CPU-reference shape, cfg3 full LM_QAGNN forward with a
random-init RoBERTa-large, cfg5 stress layer).  bench.py covers cfg2 (headline) and cfg4 (multi-GPU).
Prints one JSON object per config; meant for profiles/, not for the driver."""
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qagnn_b200  # noqa: E402
from oracle import qagnn_oracle as O  # noqa: E402  (input generators only)
from qagnn_b200 import data as Dt  # noqa: E402

dev = torch.device("cuda:0")


def gpu_time(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cfg1():
    B, n, e, D, k = 4, 50, 200, 64, 1
    inp = O.synth_graph_batch(B, n, e, D, 38, 0)
    sd = O.random_state_dict(k, D, 4, 38, "prod", 0)
    mod = qagnn_b200.QAGNN_Message_Passing(None, k, 4, 38, D, D, D).eval()
    mod.load_state_dict(sd)
    mod = mod.to(dev)
    d = {k_: v.to(dev) for k_, v in inp.items()}
    ms = gpu_time(lambda: mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"]), 50)
    mod.use_cuda_graph = True
    ms_g = gpu_time(lambda: mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"]), 200)
    torch.set_num_threads(8)
    t0 = time.perf_counter()
    for _ in range(20):
        O.message_passing_forward(sd, inp["H"], inp["edge_index"], inp["edge_type"], inp["node_type"], inp["node_score"], k, 4, 38)
    cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
    return {"config": "cfg1: 4 x 50 nodes / 200 edges, D=64, k=1", "gpu_ms_per_forward": ms, "gpu_ms_cuda_graph": ms_g,
            "cpu_oracle_ms_8_threads": cpu_ms, "edges_per_s_gpu": k * B * e / (ms_g * 1e-3)}


def cfg3():
    from transformers import RobertaConfig
    bs, nc, n, D, k, seq = 64, 5, 200, 200, 5, 100
    cfg = RobertaConfig(vocab_size=50265, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                        intermediate_size=4096, max_position_embeddings=514)
    model = qagnn_b200.LM_QAGNN(None, "roberta-large", k, 4, 38, n_concept=100000, concept_dim=D, concept_in_dim=1024,
                                n_attention_head=2, fc_dim=200, n_fc_layer=0, p_emb=0.2, p_gnn=0.2, p_fc=0.2,
                                init_range=0.02, encoder_config={"config": cfg}).eval().to(dev)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "dev.graph.adj.pk")
        recs = Dt.synth_adj_pickle(path, bs * nc, seed=0)
        for r in recs:  # concept ids must index the (reduced) embedding table
            pass
        t0 = time.perf_counter()
        cids, ntypes, scores, lens, (ei, et) = Dt.load_sparse_adj_data_with_contextnode(path, n, nc, None, use_cache=False,
                                                                                       write_cache=False)
        load_s = time.perf_counter() - t0
    cids = cids % 100000
    cids[..., 0] = 0
    packed = Dt.pack_adj(ei, et, n).to(dev)
    g = torch.Generator().manual_seed(0)
    lm = [torch.randint(3, 50000, (bs, nc, seq), generator=g).to(dev), torch.ones(bs, nc, seq, dtype=torch.long, device=dev),
          torch.zeros(bs, nc, seq, dtype=torch.long, device=dev), torch.zeros(bs, nc, seq, dtype=torch.long, device=dev)]
    dec = [cids.to(dev), ntypes.to(dev), scores.to(dev), lens.to(dev)]
    with torch.no_grad():
        full_ms = gpu_time(lambda: model(*lm, *dec, packed, None), 5, warm=2)
        flat = [x.view(bs * nc, -1) for x in lm]
        enc_ms = gpu_time(lambda: model.encoder(*flat), 5, warm=2)
        sent = model.encoder(*flat)[0]
        adj = (packed.edge_index, packed.edge_type)
        d2 = [x.view(bs * nc, *x.shape[2:]) for x in dec]
        dec_ms = gpu_time(lambda: model.decoder(sent, *d2, adj), 10, warm=3)
        dec_packed_ms = gpu_time(lambda: model.decoder(sent, *d2, packed), 10, warm=3)  # graph_ptr known: one-launch prep
        Hin = torch.randn(bs * nc, n, D, device=dev) * 0.5
        mp_ms = gpu_time(lambda: model.decoder.gnn(Hin, packed, d2[1], d2[2]), 10, warm=3)
    E = packed.edge_index.size(1)
    return {"config": "cfg3: LM_QAGNN forward, RoBERTa-large random init fp32 (24L/1024h), 64x5 x 100 tokens, loader-shaped "
                      "synthetic adj.pk, k=5, D=200", "full_forward_ms": full_ms, "encoder_ms": enc_ms, "decoder_ms": dec_ms,
            "decoder_ms_packed_batch": dec_packed_ms, "message_passing_only_ms": mp_ms, "decoder_over_message_passing": dec_packed_ms / mp_ms,
            "qa_pairs_per_s_full": bs * nc / (full_ms * 1e-3), "qa_pairs_per_s_decoder_only": bs * nc / (dec_ms * 1e-3),
            "gnn_share_of_forward": dec_ms / full_ms, "edges_in_batch": E, "loader_s_for_320_records": load_s}


def cfg5(path="auto"):
    if path == "basic":
        os.environ["QAGNN_MP_PATH"] = "basic"
    else:
        os.environ.pop("QAGNN_MP_PATH", None)
    B, n, e, D, Hh = 64, 2000, 20000, 1024, 8
    g = torch.Generator().manual_seed(0)
    N, E = B * n, B * e
    x = (torch.randn(N, D, generator=g) * 0.5).to(dev)
    extra = (torch.randn(N, D, generator=g) * 0.5).to(dev)
    ei = (torch.randint(0, n, (B, 2, e), generator=g) + (torch.arange(B) * n).view(B, 1, 1)).permute(1, 0, 2).reshape(2, E).contiguous().to(dev)
    et = torch.randint(0, 38, (E,), generator=g).to(dev)
    nt = torch.randint(0, 4, (N,), generator=g).to(dev)
    enc = torch.nn.Sequential(torch.nn.Linear(38 + 1 + 8, D), torch.nn.BatchNorm1d(D), torch.nn.ReLU(), torch.nn.Linear(D, D))
    layer = qagnn_b200.GATConvE(None, D, 4, 38, enc, head_count=Hh).eval().to(dev)
    from qagnn_b200 import _lib
    lib = _lib.load()
    prep = qagnn_b200.modeling_qagnn.GraphPrep(ei, et, nt, 4, 38, 0)
    with torch.no_grad():
        layer(x, None, None, nt, extra, prep=prep)
        lib.qagnn_profile_enable(1)
        ms = gpu_time(lambda: layer(x, None, None, nt, extra, prep=prep), 5, warm=0)
        prof = _lib.profile_read()
        lib.qagnn_profile_enable(0)
    balg = 16 * N * D + 24 * E + 8 * N
    mp_ms = prof["message_passing"][0] / max(prof["message_passing"][1], 1)
    peak = 6486.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    return {"config": "cfg5 stress: 64 x 2000 nodes / 20000 edges, D=1024, H=8, one GATConvE layer, message passing = "
                      + ("basic CSR kernels" if path == "basic" else "column-sliced kernels (edge tables in shared memory)"),
            "frac_of_measured_hbm": balg / (mp_ms * 1e-3) / 1e9 / peak, "layer_ms": ms, "message_passing_ms": mp_ms, "B_alg_bytes": balg, "mp_GBps_algorithmic": balg / (mp_ms * 1e-3) / 1e9,
            "stages_ms": {k_: v[0] / max(v[1], 1) for k_, v in prof.items() if v[1]}}


def train():
    """Training step of QAGNN_Message_Passing at cfg2 (forward with batch statistics + dropout, backward through the CUDA
    message-passing kernels, SGD step): wall time per step and the share of the hand-written kernels."""
    B, n, e, D, k = 320, 200, 1000, 200, 5
    inp = O.synth_graph_batch(B, n, e, D, 38, 100)
    sd = O.random_state_dict(k, D, 4, 38, "prod", 0)
    mod = qagnn_b200.QAGNN_Message_Passing(None, k, 4, 38, D, D, D, dropout=0.2)
    mod.load_state_dict(sd)
    mod = mod.to(dev).train()
    d = {k_: v.to(dev) for k_, v in inp.items()}
    opt = torch.optim.SGD(mod.parameters(), lr=1e-3)
    prep = mod.prepare_graph(d["edge_index"], d["edge_type"], d["node_type"].view(-1))
    prep.n_per_graph = 0

    def step():
        opt.zero_grad(set_to_none=True)
        out = mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"], prep=prep)
        out.square().mean().backward()
        opt.step()
    ms = gpu_time(step, 10, warm=3)
    from qagnn_b200 import _lib
    lib = _lib.load()
    lib.qagnn_profile_enable(1)
    step()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    lib.qagnn_profile_enable(0)
    mod.eval()
    with torch.no_grad():
        ms_eval = gpu_time(lambda: mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"]), 20)
    return {"config": "train: cfg2 batch, one optimiser step of QAGNN_Message_Passing (k=5), fp32", "ms_per_train_step": ms,
            "ms_per_eval_forward_per_kernel_launches": ms_eval, "edge_layers_per_s_training": k * B * e / (ms * 1e-3),
            "cuda_mp_forward_ms_per_step": prof["message_passing"][0]}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg3", "cfg5"]
    for name in which:
        if name.startswith("cfg5"):
            print(json.dumps({name: cfg5("basic" if name.endswith("basic") else "auto")}))
        else:
            print(json.dumps({name: globals()[name]()}))
