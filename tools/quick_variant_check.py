"""30-second in-process A/B of QAGNN_MP_VARIANT=0/1 on the cfg2 batch: bit-equality of the outputs and the
message-passing stage time of each (the library reads the variable at every launch)."""
import json, os, sys, time
t00 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "quick_variant.txt")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
def say(*a):
    with open(OUT, "a") as f:
        f.write(" ".join(str(x) for x in a) + "\n")
    print(*a, flush=True)
say("start")
import torch
say("torch imported", round(time.time() - t00, 1))
import qagnn_b200
from qagnn_b200 import _lib
from oracle import qagnn_oracle as O
lib = _lib.load()
dev = torch.device("cuda", 0)
res = {}
for name, B, n, e, D, k, realistic in (("cfg2", 320, 200, 1000, 200, 5, False), ("hubs", 32, 40, 1400, 200, 1, False), ("cfg1", 4, 50, 200, 64, 1, False)):
    inp = O.synth_graph_batch(B, n, e, D, 38, 7, realistic)
    sd = O.random_state_dict(k, D, 4, 38, "prod", 7)
    mod = qagnn_b200.QAGNN_Message_Passing(None, k, 4, 38, D, D, D).eval()
    mod.load_state_dict(sd); mod = mod.to(dev)
    d = {k_: v.to(dev) for k_, v in inp.items() if k_ != "adj_lengths"}
    outs = {}
    for variant in ("0", "1", "0", "1"):
        os.environ["QAGNN_MP_VARIANT"] = variant
        for _ in range(2):
            out = mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"])
        torch.cuda.synchronize()
        lib.qagnn_profile_enable(1)
        for _ in range(10):
            out = mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"])
        torch.cuda.synchronize()
        prof = _lib.profile_read(); lib.qagnn_profile_enable(0)
        us = prof["message_passing"][0] / 10 / k * 1e3
        outs[variant] = out.cpu()
        say(name, "variant", variant, "mp us/layer", round(us, 1), "t", round(time.time() - t00, 1))
    say(name, "bit-identical:", torch.equal(outs["0"], outs["1"]), "finite:", bool(torch.isfinite(outs["1"]).all()),
        "maxdiff", (outs["0"] - outs["1"]).abs().max().item())
say("done", round(time.time() - t00, 1))
