"""Compares the candidate consumer code of the tiled message-passing kernel (QAGNN_MP_VARIANT=1 / 2, see
qagnn_b200/csrc/mp_headtile.cu) with the default one on a B200: results must be BIT-identical (same summation trees),
and the message-passing stage time is printed for both.  The variant is chosen once per process (the library reads
the environment on its first launch), so every (variant, QPW) pair runs in a child process.

    python tools/check_mp_variant.py                 # variants 0, 1, 2 at the default warps/quad split
    python tools/check_mp_variant.py --qpw 2 3 4     # ... and with fewer, fatter warps

Exit code 0 only if every candidate output equals the default bit for bit.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # (name, graphs, nodes/graph, edges/graph, D, k, realistic)
    ("cfg1", 4, 50, 200, 64, 1, False),
    ("cfg2-uniform", 320, 200, 1000, 200, 5, False),
    ("cfg2-loader-shaped", 320, 200, 1000, 200, 5, True),
    ("hubs", 64, 40, 1400, 200, 2, False),  # 35 edges per node on average: the degree > 8 path
]


def child(out_path):
    import torch
    import qagnn_b200
    from qagnn_b200 import _lib
    from oracle import qagnn_oracle as O  # input generator only

    dev = torch.device("cuda", 0)
    lib = _lib.load()
    res, outs = {}, {}
    for name, B, n, e, D, k, realistic in CASES:
        inp = O.synth_graph_batch(B, n, e, D, 38, 7, realistic)
        sd = O.random_state_dict(k, D, 4, 38, "prod", 7)
        mod = qagnn_b200.QAGNN_Message_Passing(None, k, 4, 38, D, D, D).eval()
        mod.load_state_dict(sd)
        mod = mod.to(dev)
        d = {k_: v.to(dev) for k_, v in inp.items() if k_ != "adj_lengths"}

        def step():
            return mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"])
        for _ in range(3):
            out = step()
        torch.cuda.synchronize()
        lib.qagnn_profile_enable(1)
        steps = 20
        for _ in range(steps):
            out = step()
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        lib.qagnn_profile_enable(0)
        res[name] = {s: v[0] / steps * 1e3 / max(k, 1) for s, v in prof.items()}  # us per layer, by stage
        outs[name] = out.cpu()
    torch.save(outs, out_path + ".pt")
    with open(out_path + ".json", "w") as f:
        json.dump(res, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", default=None)
    ap.add_argument("--qpw", type=int, nargs="*", default=[0])
    args = ap.parse_args()
    if args.child:
        return child(args.child)
    import torch
    tmp = tempfile.mkdtemp()
    runs = {}
    for variant in (0, 1, 2):
        for qpw in args.qpw:
            tag = f"v{variant}_qpw{qpw}"
            env = dict(os.environ, QAGNN_MP_VARIANT=str(variant))
            if qpw:
                env["QAGNN_MP_QPW"] = str(qpw)
            path = os.path.join(tmp, tag)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path], env=env, cwd=ROOT, timeout=900)
            if r.returncode != 0:
                print(f"{tag}: child failed with exit code {r.returncode}")
                runs[tag] = None
                continue
            runs[tag] = (torch.load(path + ".pt"), json.load(open(path + ".json")))
    base_tag = f"v0_qpw{args.qpw[0]}"
    if runs.get(base_tag) is None:
        print("the default variant did not run")
        return 2
    base = runs[base_tag][0]
    ok = True
    for tag, run in runs.items():
        if run is None:
            ok = False
            continue
        outs, prof = run
        same = all(torch.equal(outs[c], base[c]) for c in base)
        ok &= same
        mp = {c: round(next((v for s, v in prof[c].items() if "message" in s or s == "mp"), float("nan")), 1) for c in prof}
        print(f"{tag}: bit-identical to {base_tag}: {same}; message-passing us/layer: {mp}")
        if not same:
            for c in base:
                if not torch.equal(outs[c], base[c]):
                    print(f"   {c}: max |diff| = {(outs[c] - base[c]).abs().max().item():.3e}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
