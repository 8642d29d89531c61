#!/usr/bin/env python
"""bench.py -- scenes/sec of a full CapNet training step (forward + loss + backward [+ gradient all-reduce]
+ Adam) on synthetic 40 000-point clouds, BASELINE.json's headline metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload at N=1 (config "c3" = BASELINE.json configs[2]): CapNet(top-down caption, relation graph, orientation
head, num_graph_steps=2, num_proposals=256, num_locals=10), batch 8 per GPU, 40 000 points of XYZ+normal+height
(7 floats), vocabulary 3 500, fp32 (TF32 off on both arms), random-init weights, synthetic scenes
(scan2cap_b200/synthetic.py).  Weak scaling: every rank gets its own 8 scenes; one NCCL all-reduce on one flat
gradient buffer per step.

One JSON line is printed by rank 0 (contract in the task statement): value = device-timed whole-job
scenes/sec with inputs resident in HBM; e2e = the same step fed from pinned HOST memory (H2D of the whole
data_dict and D2H of the loss inside the timed region); roofline = the fused ball_query+group_points kernel of
SA1 (nsample = 64), timed live with CUDA events inside the timed steps, against the measured HBM peak;
cpu_baseline = the oracle port of the reference on the host cores (bounded sample).

--impl reference runs the reference's stock code path: the UNMODIFIED lib/pointnet2 CUDA kernels re-compiled
for sm_100 (oracle/_ref/pointnet2_ref_ext.so) under the literal restatement of the reference's Python layers
(oracle/ref_model.py: unfused conv/BN/ReLU/max-pool, 256-iteration adjacency loop, per-scene graphs, per-scene
.item() syncs) -- none of our kernels or modules.  If the extension is missing it falls back to the CPU-only
oracle port on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: (batch per GPU, points, use_normal, use_multiview)
    "c3": (8, 40000, True, False),   # BASELINE.json configs[2]: XYZ+normal(+height), batch 8, 1xB200
    "c4": (4, 40000, True, True),    # BASELINE.json configs[3]: XYZ+multiview+normal, 4 scenes per GPU
    "tiny": (2, 8000, True, False),  # CI / smoke
}
VOCAB = 3500
MODEL_CFG = dict(num_proposal=256, num_locals=10, use_topdown=True, query_mode="center", graph_mode="edge_conv",
                 num_graph_steps=2, use_relation=True, use_orientation=True)
LOSS_FLAGS = dict(detection=True, caption=True, orientation=True, distance=False)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    if os.path.exists(p):
        try:
            with open(p) as f:
                pk = json.load(f)
            if float(pk["hbm_gbs"]) > 0:
                return pk, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return fallback, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(cfg_name, seed, rank):
    from scan2cap_b200 import synthetic
    B, N, use_normal, use_mv = CONFIGS[cfg_name]
    d = synthetic.make_data_dict(B, N, use_normal=use_normal, use_multiview=use_mv, num_vocabs=VOCAB,
                                 seed=seed + 100000 * rank)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in d.items()}
    num_words = int(d["lang_len"].max())
    C = host["point_clouds"].shape[-1] - 3
    return host, num_words, C


def build_model(impl, C, device):
    from scan2cap_b200 import synthetic
    from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig
    DC = ScannetDatasetConfig()
    vocab, emb, _ = synthetic.make_vocabulary(VOCAB)
    torch.manual_seed(42)
    if impl == "ours":
        from scan2cap_b200.models.capnet import CapNet
        from scan2cap_b200.lib.loss_helper import get_scene_cap_loss
    else:
        from oracle.ref_model import CapNet
        from oracle.ref_loss import get_scene_cap_loss
    model = CapNet(DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr,
                   input_feature_dim=C, **MODEL_CFG).to(device)
    model.train()
    return model, DC, get_scene_cap_loss


def to_device(host, device, num_words=None):
    d = {k: v.to(device, non_blocking=True) for k, v in host.items()}
    if num_words is not None:
        d["num_words"] = num_words  # lets the caption module skip its one D2H read (host already knows it)
    return d


def cpu_baseline(cfg_name, scenes=4, budget_s=12.0, max_steps=6):
    """Oracle port of the reference (C oracle for the native ops + restated Python layers) on the host cores."""
    from oracle import ref_model as R
    from oracle import native
    R.set_backend(None)
    _, N, use_normal, use_mv = CONFIGS[cfg_name]
    from scan2cap_b200 import synthetic
    d = synthetic.make_data_dict(scenes, N, use_normal=use_normal, use_multiview=use_mv, num_vocabs=VOCAB, seed=7)
    host = {k: torch.from_numpy(v) for k, v in d.items()}
    C = host["point_clouds"].shape[-1] - 3
    torch.set_num_threads(os.cpu_count())
    model, DC, loss_fn = build_model("reference", C, "cpu")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    steps, t0 = 0, time.perf_counter()
    while True:  # bounded sample: whole steps until >= `budget_s` seconds of CPU work (at most `max_steps`)
        out = loss_fn(model({k: v.clone() for k, v in host.items()}), "cpu", DC, None, **LOSS_FLAGS)
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or steps >= max_steps:
            break
    return {"value": scenes * steps / dt, "unit": "scenes/s", "cores": os.cpu_count(), "kind": "port",
            "threads": {"torch": torch.get_num_threads(), "openmp_native_ops": native.num_threads()},
            "sample": "%d training step(s) on %d scene(s) of the same workload (N=%d), %.1f s of host time"
                      % (steps, scenes, N, dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="ours: issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--ref-device", default="auto", choices=["auto", "cuda", "cpu"])
    args = ap.parse_args()

    # stdout carries exactly ONE JSON line: anything libraries print on fd 1 meanwhile (e.g. NCCL's version banner
    # at communicator creation) is routed to stderr, and fd 1 is restored for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    ref_on_cpu = False
    if args.impl == "reference":
        from conftest import load_reference_ext
        ext = load_reference_ext() if args.ref_device != "cpu" else None
        ref_on_cpu = ext is None or not torch.cuda.is_available()
        if ref_on_cpu:
            # CPU-only oracle port: rank 0 alone runs it
            if rank != 0:
                return
            cb = cpu_baseline(args.config, scenes=4)
            B, N, _, _ = CONFIGS[args.config]
            line = {"impl": "reference", "metric": "scenes/sec CapNet fwd+bwd @40k pts", "value": cb["value"],
                    "unit": "scenes/s", "n_gpus": 0, "steps": 1, "warmup": 0, "ms_per_step": 1e3 / cb["value"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                    "data": "synthetic", "config": {"workload": args.config, "points": N, "note": "CPU oracle port"},
                    "cpu_baseline": cb,
                    "e2e": {"value": cb["value"], "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            emit(line)
            return
        from oracle import ref_model as R
        R.set_backend(ext)

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    host, num_words, C = build_inputs(args.config, seed=42, rank=rank)
    model, DC, loss_fn = build_model(args.impl, C, device)
    B, N = host["point_clouds"].shape[0], host["point_clouds"].shape[1]
    from scan2cap_b200.distributed import FlatGradients
    engine = None
    if args.impl == "ours":
        from scan2cap_b200.engine import TrainStep
        engine = TrainStep(model, DC, lr=1e-3, weight_decay=1e-5, use_cuda_graph=not args.no_graph, **LOSS_FLAGS)
    else:
        flat = FlatGradients(model)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)

    # make the referred box of every scene a box the (random-init) detector proposes, so good_bbox_masks is not
    # empty and the caption loss / its gradients are exercised
    with torch.no_grad():
        probe = model(to_device(host, device, num_words))
        host["ref_box_corner_label"] = probe["bbox_corner"][:, 3].detach().cpu().pin_memory()
    del probe

    import scan2cap_b200._lib as L
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # 256 MB > 126 MB L2

    def step(data):
        if engine is not None:
            return engine.run(data)
        flat.zero_()
        out = loss_fn(model(data), device, DC, None, **LOSS_FLAGS)
        out["loss"].backward()
        flat.all_reduce_mean()
        opt.step()
        return out["loss"]

    def timed(K, from_host):
        resident = None if from_host else to_device(host, device, num_words)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        last = None
        nxt = None
        if from_host and engine is not None:
            # double-buffered input pipeline of the public API: every step's inputs travel pinned host -> device inside
            # the timed region (K+1 transfers for K steps), each one overlapping the previous step's compute
            nxt = dict(host, num_words=num_words)
            engine.prefetch(nxt)
        for _ in range(K):
            flush.zero_()  # L2 flush between iterations (inside the timed region: conservative)
            if from_host:
                if engine is not None:  # the engine copies pinned host tensors into its static device buffers
                    cur = nxt
                    loss = step(cur)
                    nxt = dict(host, num_words=num_words)
                    engine.prefetch(nxt)  # next step's H2D, in flight while this step computes
                else:
                    loss = step(to_device(host, device, None))
                last = float(loss.item())  # D2H read of the step's result
            else:
                loss = step({k: v for k, v in resident.items()})
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    timed(args.warmup, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.LAUNCH_COUNT = 0
    ms, _ = timed(args.steps, False)
    launches = L.LAUNCH_COUNT
    if engine is not None and engine.use_graph:
        launches = engine.kernels_per_step * args.steps  # replayed from the graph: no Python call per launch
    clocks = sampler.stop() if rank == 0 else None
    # single-kernel timing for the roofline: the same step issued eagerly, CUDA events around the SA1
    # query+group launch (events cannot bracket a node inside a replayed graph)
    qg_name, qg_grid = "s2c_query_and_group", "s2c_query_and_group_grid"
    qg_events, qg_steps, qg_kernel = [], min(args.steps, 5), None
    if engine is not None:
        resident = to_device(host, device, num_words)
        L.TIMING = {qg_name: [], qg_grid: []}
        for _ in range(qg_steps):
            flush.zero_()
            engine.run_eager(dict(resident))
        torch.cuda.synchronize()
        if L.TIMING[qg_grid]:
            # SA1 (n = 40 000 >= S2C_BALL_GRID_MIN) is the only call of the step that takes the uniform-grid entry
            # point: grid_build_kernel + grid_query_kernel<GROUP>, both inside the timed bracket
            per = len(L.TIMING[qg_grid]) // qg_steps
            qg_events = [e for i, e in enumerate(L.TIMING[qg_grid]) if i % per == 0]
            qg_kernel = "grid_build_kernel + grid_query_kernel<GROUP>"
        else:
            # SA1 is the first query_and_group call of every step (5 per step: SA1-4 + vote aggregation)
            per = len(L.TIMING[qg_name]) // qg_steps
            qg_events = [e for i, e in enumerate(L.TIMING[qg_name]) if i % per == 0]
            qg_kernel = "ball_query_kernel<GROUP>"
        L.TIMING = None
    timed(2, True)
    ms_e2e, last_loss = timed(args.steps, True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    h2d = int(sum(v.numel() * v.element_size() for v in host.values()))
    scenes = B * world * args.steps
    pk, pk_src = peaks()
    line = {
        "metric": "scenes/sec CapNet fwd+bwd @40k pts", "value": scenes / (ms * 1e-3), "unit": "scenes/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: full CapNet training step (zero_grad+fwd+loss+bwd+grad all-reduce+Adam), "
                               "batch %d/GPU, %d pts x %d floats, K=256 proposals, L=10 locals, 2 graph steps, "
                               "top-down caption, V=%d, %d decoder steps" % (args.config, B, N, C + 3, VOCAB, num_words - 1),
                   "global_batch": B * world, "points": N, "point_floats": C + 3, "tf32": False,
                   "l2": "256 MB buffer rewritten between iterations (inside the timed region)",
                   "parallelism": "dp%d" % world,
                   "issue": ("cuda-graph replay of the whole step" if (engine is not None and engine.use_graph)
                             else "eager launches")},
        "e2e": {"value": scenes / (ms_e2e * 1e-3), "unit": "scenes/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "last_loss": last_loss},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if args.impl == "reference":
        line["impl"] = "reference"
        line["gpu_launches"] = 0
        line["config"]["reference"] = ("unmodified lib/pointnet2 CUDA kernels (sm_100 build, oracle/_ref) + literal "
                                       "restatement of the reference Python layers; CUDA_LAUNCH_BLOCKING unset")
    if qg_events:
        t_ms = float(np.mean([s.elapsed_time(e) for s, e in qg_events]))
        M, ns = 2048, 64
        alg = B * (12 * N + 12 * M + 4 * C * N + 4 * M * ns + 4 * (3 + C) * M * ns)
        achieved = alg / (t_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(args.config, {}).get("traffic")
        line["roofline"] = {"bound": "hbm", "kernel": "%s (SA1: N=%d, M=2048, nsample=64, C=%d)" % (qg_kernel, N, C),
                            "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                            "traffic": traffic, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": t_ms,
                            "peak_source": pk_src}
    if not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline(args.config, scenes=4)
        except Exception as e:  # never lose the GPU numbers to a host-side problem
            line["cpu_baseline"] = {"error": repr(e)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
