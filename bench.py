#!/usr/bin/env python
"""bench.py — the QA-GNN message-passing hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One "step" = one QAGNN_Message_Passing.forward (k=5 GATConvE layers, modeling_qagnn.py:53-95) over one
synthetic batch of BASELINE.json configs[1]: 64x5 = 320 sub-graphs of 200 nodes / 1000 edges per GPU,
hidden 200, 4 heads, 38 relation types.  Metric: GNN edges/sec = k * E / t ("edge-layers per second",
E = real directed edges, self loops excluded).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(graphs=320, n=200, e=1000, D=200, k=5, H=4, T=4, R=38)
METRIC = "GNN edges/sec"
UNIT = "edge-layers/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-graphs", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch the ~60 kernels of a step one by one")
    return ap.parse_args()


def workload_name():
    return (f"cfg2: {CFG['graphs']} graphs/GPU (64x5 choices) x {CFG['n']} nodes x {CFG['e']} edges, hidden {CFG['D']}, "
            f"{CFG['k']} layers, {CFG['H']} heads, {CFG['R']} edge types (17 merged relations -> (17+2)*2)")


def b_alg_per_layer(N, E, D):
    """compulsory bytes of the message-passing kernel per layer (read Q,Kx,Mx + write aggr,
    int64 src/tgt/etype, int64 node_type)."""
    return 16 * N * D + 24 * E + 8 * N


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (oracle/qagnn_oracle.py), all host threads
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_graphs, seed=0, budget_s=None):
    """Times the CPU restatement on `sample_graphs` graphs of the workload.  With `budget_s` the sample is shrunk
    (never below 4 graphs) so that warmup+steps forwards fit the budget; steps and warmup are always honoured."""
    from oracle import qagnn_oracle as O
    ncpu = os.cpu_count() or 1
    sd = O.random_state_dict(CFG["k"], CFG["D"], CFG["T"], CFG["R"], "prod", seed)
    # torch's intra-op pool degrades badly when oversubscribed on many-core hosts: probe a few pool sizes on a
    # small slice and keep the fastest ("all the host threads it can use")
    probe = O.synth_graph_batch(8, CFG["n"], CFG["e"], CFG["D"], CFG["R"], seed)
    best = (float("inf"), 1)
    for nt in sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu}):
        torch.set_num_threads(nt)
        O.message_passing_forward(sd, probe["H"], probe["edge_index"], probe["edge_type"], probe["node_type"],
                                  probe["node_score"], 1, CFG["T"], CFG["R"], CFG["H"])
        t0 = time.perf_counter()
        O.message_passing_forward(sd, probe["H"], probe["edge_index"], probe["edge_type"], probe["node_type"],
                                  probe["node_score"], 1, CFG["T"], CFG["R"], CFG["H"])
        best = min(best, (time.perf_counter() - t0, nt))
    cores = best[1]
    torch.set_num_threads(cores)
    if budget_s is not None:
        per_graph = best[0] / 8 * CFG["k"]          # probe = 8 graphs, 1 layer
        sample_graphs = int(max(4, min(sample_graphs, budget_s / ((steps + warmup) * per_graph))))
    inp = O.synth_graph_batch(sample_graphs, CFG["n"], CFG["e"], CFG["D"], CFG["R"], seed)
    E = inp["edge_index"].size(1)

    def step():
        return O.message_passing_forward(sd, inp["H"], inp["edge_index"], inp["edge_type"], inp["node_type"],
                                         inp["node_score"], CFG["k"], CFG["T"], CFG["R"], CFG["H"])
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": CFG["k"] * E / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_graphs} of the {CFG['graphs']} graphs of the workload ({E} edges), {steps} timed "
                      f"forwards of the op-for-op oracle port (torch CPU fp32, {cores} threads = fastest pool size of those probed on "
                      f"this {ncpu}-cpu host), {dt * 1e3:.1f} ms each",
            "ms_per_step": dt * 1e3}


def gemm_roofline(N, D, avg_launch_ms, peaks):
    """Tensor roofline of the projection GEMM [N,2D]x[2D,3D]: three bf16 passes (hi*hi, hi*lo, lo*hi) per launch."""
    peak = peaks.get("bf16_tflops", 1590.0)
    flops = 3 * 2 * N * (2 * D) * (3 * D)
    achieved = flops / (avg_launch_ms * 1e-3) / 1e12 if avg_launch_ms > 0 else 0.0
    return {"kernel": "gemm_tc_kernel (projection Q|Kx|Mx)", "bound": "tensor", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak, "avg_launch_ms": avg_launch_ms}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    r = cpu_reference_run(steps, warm, args.cpu_sample_graphs, budget_s=100.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "bounded_sample": r["sample"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference's torch-geometric/torch-scatter wheels cannot be installed offline; this arm times the "
                "op-for-op CPU restatement of the reference path (oracle/), pinned against goldens minted from the "
                "reference's own modules",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()  # exact child we started
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch.distributed as dist
    import qagnn_b200
    from qagnn_b200 import _lib
    from oracle import qagnn_oracle as O  # input generator only (shared with the tests)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B, n, e, D, k = CFG["graphs"], CFG["n"], CFG["e"], CFG["D"], CFG["k"]
    inp = O.synth_graph_batch(B, n, e, D, CFG["R"], seed=100 + rank)
    sd = O.random_state_dict(k, D, CFG["T"], CFG["R"], "prod", seed=0)
    mod = qagnn_b200.QAGNN_Message_Passing(None, k, CFG["T"], CFG["R"], D, D, D).eval()
    mod.load_state_dict(sd)
    mod = mod.to(dev)
    N, E = B * n, inp["edge_index"].size(1)

    host = {k_: v.pin_memory() for k_, v in inp.items() if k_ != "adj_lengths"}
    d = {k_: v.to(dev, non_blocking=True) for k_, v in host.items()}
    out_host = torch.empty(B, n, D, dtype=torch.float32).pin_memory()
    gathered = torch.empty(world * B, D, device=dev) if world > 1 else None

    # resident arm: graph prep + all layers replayed as ONE CUDA graph (inputs stay in the same device buffers);
    # indices are validated once before the capture.  Falls back to per-kernel launches if capture is unavailable.
    launch_mode = "per-kernel launches"
    if not args.no_cuda_graph:
        try:
            mod.use_cuda_graph = True
            mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"])
            torch.cuda.synchronize()
            launch_mode = "one CUDA graph per step (prep + 5 layers + epilogue)"
        except Exception as exc:  # noqa: BLE001
            mod.use_cuda_graph = False
            mod._graphs.clear()
            launch_mode = f"per-kernel launches (CUDA graph capture failed: {type(exc).__name__})"

    def step_resident():
        out = mod(d["H"], (d["edge_index"], d["edge_type"]), d["node_type"], d["node_score"])
        if world > 1:  # the path's single collective: all-gather of the pooled (context-node) vectors
            dist.all_gather_into_tensor(gathered, out[:, 0].contiguous())
        return out

    # e2e arm: the public streaming API (qagnn_b200.pipeline.StreamedRunner): every step uploads its inputs from pinned
    # host memory, runs the forward and downloads the [B,n,D] result; copies run on their own streams so that step i+1's
    # H2D and step i-1's D2H overlap step i's kernels (two device buffer sets, one captured CUDA graph each)
    from qagnn_b200.pipeline import StreamedRunner
    runner = StreamedRunner(mod, host, dev, depth=2)

    def step_e2e():
        slot = runner.submit(host)
        if world > 1:  # the path's single collective, on the compute stream right after the forward
            with torch.cuda.stream(runner.compute):
                dist.all_gather_into_tensor(gathered, runner.dev_out[slot][:, 0].contiguous())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        if profile:
            lib.qagnn_profile_enable(1)
        l0 = lib.qagnn_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = lib.qagnn_launch_count() - l0
        prof = _lib.profile_read() if profile else None
        if profile:
            lib.qagnn_profile_enable(0)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), launches, prof

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, _, _ = timed(step_resident, args.steps)
    for _ in range(4):
        step_e2e()
    runner.drain()

    def e2e_loop():
        step_e2e()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s_ in (runner.h2d, runner.compute, runner.d2h):
        s_.wait_event(ev0)
    for _ in range(args.steps):
        e2e_loop()
    runner.drain()
    ev1.record()
    barrier()
    t_ = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
    ms_e2e = t_.item()
    # kernel-level pass: the same K steps launched kernel by kernel with the library's CUDA-event stage timers on the
    # launching stream (events cannot be timed inside a replayed graph); feeds `roofline`, `stages`, `gpu_launches`
    graphed = mod.use_cuda_graph
    mod.use_cuda_graph = False
    step_resident()
    ms_eager, launches, prof = timed(step_resident, args.steps, profile=True)
    mod.use_cuda_graph = graphed
    clocks = sampler.stop() if rank == 0 else None

    ms_step = ms_total / args.steps
    value = world * k * E / (ms_step * 1e-3)
    e2e_value = world * k * E / (ms_e2e / args.steps * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * 4

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    mp_ms, mp_cnt = prof["message_passing"]
    mp_avg_ms = mp_ms / max(mp_cnt, 1)
    balg = b_alg_per_layer(N, E, D)
    achieved = balg / (mp_avg_ms * 1e-3) / 1e9 if mp_avg_ms > 0 else 0.0
    stages = {s: {"ms_per_step": v[0] / args.steps, "intervals_per_step": v[1] / args.steps} for s, v in prof.items()}
    stage_total = sum(v["ms_per_step"] for v in stages.values()) or 1.0
    for v in stages.values():
        v["share"] = v["ms_per_step"] / stage_total

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(), "graphs_per_gpu": B, "N": N, "E": E, "parallelism": f"dp{world} (sub-graph "
                   "shards, one NCCL all-gather of pooled vectors)" if world > 1 else "single GPU",
                   "l2": f"no flush: a step streams the {lib.qagnn_forward_workspace_bytes(_lib.C.byref(mod._shape(N, E, n))) / 1e6:.0f} MB "
                         "workspace + 51 MB inputs, > 126 MB L2", "step": "graph prep + 5 x (projection, message passing, "
                   "node MLP) + Vh/Vx epilogue, inputs resident in HBM", "launch": launch_mode},
        "qa_pairs_per_s": world * B / (ms_step * 1e-3),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps,
                "api": "qagnn_b200.pipeline.StreamedRunner around QAGNN_Message_Passing.forward: per step H2D of H/edge_index/"
                       "edge_type/node_type/node_score from pinned host memory, forward, D2H of the [B,n,D] output into pinned "
                       "memory; copies on separate streams, double-buffered, so they overlap the neighbouring steps' kernels"},
        "gpu_launches": int(launches),
        "gpu_launches_note": "kernels of libqagnn_b200.so enqueued by the K steps of the kernel-level pass (the CUDA graph "
                             "of the headline pass replays the same kernel nodes)",
        "ms_per_step_per_kernel_launches": ms_eager / args.steps,
        "roofline": {"kernel": "mp_headtile_kernel (fused logits + per-source softmax + out-degree rescale + per-target sum), one GATConvE layer",
                     "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": balg, "avg_launch_ms": mp_avg_ms,
                     "launches_timed": int(mp_cnt),
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r1_mp_headtile_ncu.md)
                     "traffic": 191374336 if world == 1 else None},  # 173.0 MB read + 18.3 MB written
        # second-largest kernel: the tcgen05 projection GEMM [N,2D]x[2D,3D], three bf16 passes (hi*hi, hi*lo, lo*hi)
        "roofline_gemm": gemm_roofline(N, D, prof["projection"][0] / max(prof["projection"][1], 1), peaks),
        "stages": stages,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:  # the CPU arm is reported at N=1 only
        r = cpu_reference_run(3, 1, args.cpu_sample_graphs)
        line["cpu_baseline"] = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
