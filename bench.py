#!/usr/bin/env python
"""bench.py -- scenes/sec of a full CapNet training step (forward + loss + backward [+ gradient all-reduce]
+ Adam) on synthetic 40 000-point clouds, BASELINE.json's headline metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Headline workload (all N; config "c3" = BASELINE.json configs[2]): CapNet(top-down caption, relation graph,
orientation head, num_graph_steps=2, num_proposals=256, num_locals=10), batch 8 per GPU, 40 000 points of
XYZ+normal+height (7 floats), vocabulary 3 500, fp32 (TF32 off on both arms), random-init weights, synthetic scenes
(scan2cap_b200/synthetic.py).  Weak scaling: every rank gets its own 8 scenes; one NCCL all-reduce on one flat
gradient buffer per step, captured inside the step's CUDA graph.

The same run also measures BASELINE.json configs[3] ("config4" in the JSON line: 4 scenes per GPU of
XYZ+multiview+normal+height = 135 floats per point, the configuration the 8-GPU target is quoted on) with the same
timing rules, and the fused ball_query+group_points kernel at the shape where bytes dominate ("roofline_c132").

One JSON line is printed by rank 0 (contract in the task statement): value = device-timed whole-job
scenes/sec with inputs resident in HBM; e2e = the same step fed from pinned HOST memory (H2D of the whole
data_dict and D2H of the loss inside the timed region); roofline = the fused ball_query+group_points kernel of
SA1 (nsample = 64), timed live with CUDA events inside the timed steps, against the measured HBM peak;
cpu_baseline = BASELINE configs[0], the reference's own CPU-runnable capnet_pretrained path on the host cores.

--impl reference runs the reference's stock code path and NOTHING of ours: the UNMODIFIED reference Python modules
(baseline/_ref, a git-ignored verbatim copy made by baseline/install_ref.py) over the UNMODIFIED lib/pointnet2
CUDA kernels re-compiled for sm_100 (oracle/_ref/pointnet2_ref_ext.so), issued as lib/solver.py issues a training
iteration.  That process imports neither scan2cap_b200 nor the oracle restatements (asserted before the line is
printed).  With N > 1 each rank is an independent replica (the reference has no multi-GPU mode).
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
METRIC = "scenes/sec CapNet fwd+bwd @40k pts"


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


def trace(msg):
    """Progress marker on stderr (S2C_BENCH_TRACE=1): where a multi-rank run stops if it ever hangs."""
    if os.environ.get("S2C_BENCH_TRACE"):
        sys.stderr.write("[bench rank %s %.1fs] %s\n" % (os.environ.get("RANK", "0"), time.time() - _T0, msg))
        sys.stderr.flush()


_T0 = time.time()


def finish(world):
    """Leave without tearing NCCL down: destroy_process_group() blocks while captured graphs still hold the
    communicator's kernels (observed: the JSON line printed after 11 s, the process never exited).  Every rank has
    passed the final barrier of the timed region by now; flush and exit 0."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def workload_string(name, B, N, F, num_words):
    return ("%s: full CapNet training step (zero_grad+fwd+loss+bwd+grad all-reduce+Adam), batch %d/GPU, %d pts x %d "
            "floats, K=256 proposals, L=10 locals, 2 graph steps, top-down caption, V=%d, caption length %d words"
            % (name, B, N, F, VOCAB, num_words))


def cpu_baseline(budget_s=12.0):
    """BASELINE configs[0] on the host cores, in a SUBPROCESS (its shims neutralise Tensor.cuda process-wide)."""
    cmd = [sys.executable, "-m", "baseline.reference_arm", "--cpu-pretrained", "--budget", str(budget_s)]
    try:
        p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"error": (p.stderr or p.stdout)[-400:]}
    except Exception as e:  # never lose the GPU numbers to a host-side problem
        return {"error": repr(e)}


class Timer(object):
    """K steps bracketed by barrier + synchronize on both sides, CUDA events, max over ranks."""

    def __init__(self, device, world):
        self.device, self.world = device, world
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # 256 MB > 126 MB L2

    def run(self, K, body):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            self.flush.zero_()  # L2 flush between iterations (inside the timed region: conservative)
            body(i)
        b.record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=self.device)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class Ours(object):
    def __init__(self, cfg_name, device, rank, use_graph=True):
        from scan2cap_b200 import synthetic
        from scan2cap_b200.data.scannet.model_util_scannet import ScannetDatasetConfig
        from scan2cap_b200.engine import TrainStep
        from scan2cap_b200.models.capnet import CapNet
        self.name, self.device = cfg_name, device
        B, N, use_normal, use_mv = CONFIGS[cfg_name]
        d = synthetic.make_data_dict(B, N, use_normal=use_normal, use_multiview=use_mv, num_vocabs=VOCAB,
                                     seed=42 + 100000 * rank)
        self.host = {k: torch.from_numpy(v).pin_memory() for k, v in d.items()}
        self.num_words = int(d["lang_len"].max())
        self.B, self.N, self.F = B, N, self.host["point_clouds"].shape[-1]
        DC = ScannetDatasetConfig()
        vocab, emb, _ = synthetic.make_vocabulary(VOCAB)
        torch.manual_seed(42)
        self.model = CapNet(DC.num_class, vocab, emb, DC.num_heading_bin, DC.num_size_cluster, DC.mean_size_arr,
                            input_feature_dim=self.F - 3, **MODEL_CFG).to(device)
        self.model.train()
        self.engine = TrainStep(self.model, DC, lr=1e-3, weight_decay=1e-5, use_cuda_graph=use_graph, **LOSS_FLAGS)
        # make the referred box of every scene a box the (random-init) detector proposes, so good_bbox_masks is not
        # empty and the caption loss / its gradients are exercised
        with torch.no_grad():
            probe = self.model(self.resident())
            self.host["ref_box_corner_label"] = probe["bbox_corner"][:, 3].detach().cpu().pin_memory()
        del probe

    def resident(self):
        d = {k: v.to(self.device, non_blocking=True) for k, v in self.host.items()}
        d["num_words"] = self.num_words
        return d

    def h2d_bytes(self):
        return int(sum(v.numel() * v.element_size() for v in self.host.values()))

    def timed(self, timer, K, from_host):
        """K steps through the public pipeline API of TrainStep: prefetch(next batch) -- input copies (from pinned HOST
        memory when from_host, else device-to-device from resident tensors) and the coordinate-only FPS indices of the
        next batch, on a side stream -- overlapping run(current batch).  K steps issue K prefetches inside the timed
        region (the first batch's is issued before it, the (K+1)-th inside)."""
        eng = self.engine
        src = self.host if from_host else {k: v for k, v in self.resident().items() if isinstance(v, torch.Tensor)}
        torch.cuda.synchronize()   # resident sources are complete before the copy stream reads them
        state = {"nxt": dict(src, num_words=self.num_words), "last": None}
        eng.prefetch(state["nxt"])

        def body(i):
            cur = state["nxt"]
            loss = eng.run(cur)
            state["nxt"] = dict(src, num_words=self.num_words)
            eng.prefetch(state["nxt"])            # next step's inputs + indices, in flight while this step computes
            if from_host:
                state["last"] = float(loss.item())    # D2H read of the step's result
        ms = timer.run(K, body)
        return ms, state["last"]

    def query_group_events(self, timer, steps):
        """CUDA-event time of the SA1 fused query+group entry point inside eager steps (events cannot bracket a node
        of a replayed graph).  SA1 (n = 40 000) is the only call of the step that takes the uniform-grid entry."""
        import scan2cap_b200._lib as L
        res = self.resident()
        names = ("s2c_query_and_group_grid_prebuilt", "s2c_query_and_group_grid", "s2c_query_and_group")
        for _ in range(2):   # untimed: the first eager steps of a process that has only replayed graphs so far pay lazy
            timer.flush.zero_()   # initialisation (allocator growth, module loading) inside the bracketed call
            self.engine.run_eager(dict(res))
        torch.cuda.synchronize()
        L.TIMING = {n: [] for n in names}
        for _ in range(steps):
            timer.flush.zero_()
            self.engine.run_eager(dict(res))
        torch.cuda.synchronize()
        ev, kernel = [], None
        for n, k in zip(names, ("group_rows_tma_kernel / grid_query_kernel<GROUP> (grid built ahead by TrainStep.prefetch)",
                                "grid_build_kernel + grid_query_kernel<GROUP>", "ball_query_kernel<GROUP>")):
            if L.TIMING[n]:
                per = len(L.TIMING[n]) // steps
                ev = [e for i, e in enumerate(L.TIMING[n]) if i % per == 0]  # SA1 = the first call of every step
                kernel = k
                break
        L.TIMING = None
        return [s.elapsed_time(e) for s, e in ev], kernel


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    kernel at this shape (profiles/roofline_traffic.json names the .txt summary it was read from); None if absent."""
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(tp) as f:
            return json.load(f).get(key, {}).get("traffic")
    except Exception:
        return None


def qg_roofline(t_ms, kernel, B, N, C, pk, pk_src, layout, traffic=None):
    M, ns = 2048, 64
    alg = B * (12 * N + 12 * M + 4 * C * N + 4 * M * ns + 4 * (3 + C) * M * ns)
    achieved = alg / (t_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "%s (SA1: B=%d, N=%d, M=2048, nsample=64, C=%d; %s)" % (kernel, B, N, C, layout),
            "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "traffic": traffic, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": t_ms, "peak_source": pk_src}


def standalone_qg(timer, B, N, C, iters=20):
    """The fused entry point on the ALIGNED feature layout (features (B,N,C) contiguous), timed alone with the L2
    flushed before every launch: the BASELINE configs[4] sweep row N=40k / C=132 / nsample=64."""
    from scan2cap_b200 import synthetic
    from scan2cap_b200.lib.pointnet2 import _ext
    pc, _ = synthetic.make_point_clouds(B, N, use_normal=False, use_height=False, seed=42)
    xyz = torch.from_numpy(np.ascontiguousarray(pc[..., :3])).cuda()
    _, new_xyz = _ext.furthest_point_sampling_with_xyz(xyz, 2048)
    feats = torch.randn((B, N, C), device="cuda", generator=torch.Generator(device="cuda").manual_seed(N + C))

    def fn():
        return _ext.query_and_group(xyz, new_xyz, feats, 0.2, 64, True, feat_point_major=True, channels_last=True,
                                    pad4=True)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        timer.flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


def run_ours(args, emit, rank, world, local_rank):
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import scan2cap_b200._lib as L
    timer = Timer(device, world)
    pk, pk_src = peaks()
    trace("process group up")
    main = Ours(args.config, device, rank, use_graph=not args.no_graph)
    trace("model built")
    predict = None
    if not args.no_predict and rank == 0:
        # benchmark/predict.py path: eval forward with greedy decoding of all 256 proposals, device-side
        # NMS, per-scene output lists; wall clock per batch including the device->host transfer and host assembly
        from scan2cap_b200.lib.predict import CaptionPredictor
        vocab = main.model.caption.vocabulary
        pred = CaptionPredictor(main.model, main.engine.DC, vocab)
        batch = {k: v for k, v in main.resident().items() if isinstance(v, torch.Tensor)}
        for _ in range(2):
            pred.predict_batch(dict(batch))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            outp = pred.predict_batch(dict(batch))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        predict = {"value": main.B / dt, "unit": "scenes/s", "ms_per_batch": 1e3 * dt, "batch": main.B,
                   "captions_per_batch": int(sum(len(v) for v in outp.values())),
                   "what": "benchmark/predict.py:170-227 for one batch of %d scenes: eval forward (graph replay) + 256 x 29 "
                           "greedy decode + device NMS + D2H + per-scene caption lists" % main.B}
        main.model.train()
        trace("predict done (freshly initialised weights: before any training step)")

    main.timed(timer, args.warmup, False)
    trace("warm-up done (graph captured)")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.LAUNCH_COUNT = 0
    ms, _ = main.timed(timer, args.steps, False)
    launches = L.LAUNCH_COUNT
    if main.engine.use_graph:
        launches = main.engine.kernels_per_step * args.steps  # replayed from the graph: no Python call per launch
    clocks = sampler.stop() if rank == 0 else None
    trace("timed steps done")
    qg_ms, qg_kernel = main.query_group_events(timer, min(args.steps, 5))
    trace("eager kernel timing done")
    main.timed(timer, 2, True)
    ms_e2e, last_loss = main.timed(timer, args.steps, True)
    trace("e2e done")

    second = None
    if args.config == "c3" and not args.no_config4:
        # BASELINE configs[3] in the same run, same rules (fewer steps: it is the secondary line)
        c4 = Ours("c4", device, rank, use_graph=not args.no_graph)
        k4 = max(3, min(args.steps, 10))
        c4.timed(timer, max(3, min(args.warmup, 5)), False)
        ms4, _ = c4.timed(timer, k4, False)
        qg4_ms, qg4_kernel = c4.query_group_events(timer, 5)
        c4.timed(timer, 2, True)
        ms4_e2e, loss4 = c4.timed(timer, k4, True)
        second = {
            "workload": workload_string("c4", c4.B, c4.N, c4.F, c4.num_words), "n_gpus": world,
            "global_batch": c4.B * world, "steps": k4, "value": c4.B * world * k4 / (ms4 * 1e-3), "unit": "scenes/s",
            "ms_per_step": ms4 / k4,
            "e2e": {"value": c4.B * world * k4 / (ms4_e2e * 1e-3), "unit": "scenes/s", "ms_per_step": ms4_e2e / k4,
                    "h2d_bytes_per_step": c4.h2d_bytes(), "d2h_bytes_per_step": 4, "last_loss": loss4},
            "gpu_launches_per_step": c4.engine.kernels_per_step}
        rc132 = {}
        if qg4_ms:
            rc132["product_layout"] = qg_roofline(
                float(np.mean(qg4_ms)), qg4_kernel, c4.B, c4.N, c4.F - 3, pk, pk_src,
                "inside the c4 training step; features = columns 3.. of the engine's 16-byte aligned point_clouds rows",
                traffic=ncu_traffic("c4"))
        if rank == 0 and world == 1:
            t = standalone_qg(timer, 8, 40000, 132)
            rc132["aligned_layout"] = qg_roofline(t, "grid_build_kernel + grid_query_kernel<GROUP>", 8, 40000, 132, pk,
                                                  pk_src, "standalone launch, L2 flushed; features (B,N,132) contiguous",
                                                  traffic=ncu_traffic("sweep_n40k_c132_ns64"))
        del c4

    if rank != 0:
        finish(world)
        return
    scenes = main.B * world * args.steps
    line = {
        "metric": METRIC, "value": scenes / (ms * 1e-3), "unit": "scenes/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.config, main.B, main.N, main.F, main.num_words),
                   "global_batch": main.B * world, "points": main.N, "point_floats": main.F, "tf32": False,
                   "l2": "256 MB buffer rewritten between iterations (inside the timed region)",
                   "parallelism": "dp%d" % world},
        "issue": ("cuda-graph replay of the whole step (%d decoder steps: caption length rounded up to a multiple of 4)"
                  % (main.engine._words({"num_words": main.num_words, "lang_ids": main.host["lang_ids"]}) - 1)
                  if main.engine.use_graph else "eager launches"),
        "e2e": {"value": scenes / (ms_e2e * 1e-3), "unit": "scenes/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": main.h2d_bytes(), "d2h_bytes_per_step": 4, "last_loss": last_loss},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    # roofline of the kernel BASELINE.json's metric names (fused ball_query + group_points, nsample = 64).  The primary
    # block is the shape where its bytes dominate -- BASELINE configs[3] (C = 132), timed inside the c4 training step of
    # this same run; at the c3 shape (C = 4: 5 MB per scene, byte floor 6.5 us) the kernel is launch / latency bound and
    # is reported as roofline_c3.
    r_c3 = None
    if qg_ms:
        r_c3 = qg_roofline(float(np.mean(qg_ms)), qg_kernel, main.B, main.N, main.F - 3, pk, pk_src,
                           "inside the c3 training step", traffic=ncu_traffic(args.config))
    if second is not None:
        line["config4"] = second
        if rc132.get("product_layout"):
            line["roofline"] = rc132["product_layout"]
            if r_c3:
                line["roofline_c3"] = r_c3
            if rc132.get("aligned_layout"):
                line["roofline_standalone"] = rc132["aligned_layout"]
        elif r_c3:
            line["roofline"] = r_c3
    elif r_c3:
        line["roofline"] = r_c3
    if predict is not None:
        line["predict"] = predict
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    finish(world)


# ------------------------------------------------------------------------------------------------------------------
# reference arm: nothing of ours in this process
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, emit, rank, world, local_rank):
    from baseline import reference_arm as RA
    from baseline import shims
    have_gpu = torch.cuda.is_available() and args.ref_device != "cpu"
    if not have_gpu or shims.load_reference_ext() is None:
        # no GPU / extension: the reference's CPU-runnable case (BASELINE configs[0]); rank 0 alone runs it
        if rank != 0:
            return
        cb = RA.time_cpu_pretrained()
        RA.assert_clean_process()
        emit({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "scenes/s", "n_gpus": 0,
              "steps": 1, "warmup": 0, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": cb["config"]},
              "cpu_baseline": cb,
              "e2e": {"value": cb["value"], "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)  # timing barrier / max over ranks only: no data-path collective
    timer = Timer(device, world)

    def measure(cfg_name, K, W):
        B, N, use_normal, use_mv = CONFIGS[cfg_name]
        syn = RA.load_synthetic()
        F = 3 + 3 * use_normal + 128 * use_mv + 1
        ref = RA.GpuReference(F - 3, VOCAB, device)
        d = syn.make_data_dict(B, N, use_normal=use_normal, use_multiview=use_mv, num_vocabs=VOCAB,
                               seed=42 + 100000 * rank, mean_size_arr=ref.DC.mean_size_arr)
        host = {k: torch.from_numpy(v).pin_memory() for k, v in d.items()}
        with torch.no_grad():
            probe = ref.forward({k: v.to(device) for k, v in host.items()})
            host["ref_box_corner_label"] = probe["bbox_corner"][:, 3].detach().cpu().pin_memory()
        del probe
        resident = {k: v.to(device) for k, v in host.items()}
        predict = None
        if cfg_name == args.config and not args.no_predict and rank == 0:
            # the stock predict loop on a BOUNDED sample (2 scenes: its per-token host loop takes seconds per batch)
            nb = min(2, B)
            small = {k: v[:nb].clone() for k, v in resident.items()}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            outp = ref.predict_batch(dict(small), ref.vocabulary)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            predict = {"value": nb / dt, "unit": "scenes/s", "ms_per_batch": 1e3 * dt, "batch": nb,
                       "captions_per_batch": int(sum(len(v) for v in outp.values())),
                       "what": "benchmark/predict.py:170-227 for one batch of %d scenes (bounded sample), stock code" % nb}
        timer.run(W, lambda i: ref.step(dict(resident)))
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ms = timer.run(K, lambda i: ref.step(dict(resident)))
        clocks = sampler.stop() if rank == 0 else None
        last = [None]

        def body(i):
            last[0] = float(ref.step(dict(host)).item())   # solver.py:380-382 moves every key .to(device); D2H of the loss
        timer.run(2, body)
        ms_e2e = timer.run(K, body)
        h2d = int(sum(v.numel() * v.element_size() for v in host.values()))
        return dict(B=B, N=N, F=F, ms=ms, ms_e2e=ms_e2e, clocks=clocks, h2d=h2d, last=last[0],
                    num_words=int(d["lang_len"].max()), predict=predict)

    m = measure(args.config, args.steps, args.warmup)
    second = None
    if args.config == "c3" and not args.no_config4:
        k4 = max(3, min(args.steps, 10))
        m4 = measure("c4", k4, 3)
        second = {"workload": workload_string("c4", m4["B"], m4["N"], m4["F"], m4["num_words"]), "n_gpus": world,
                  "global_batch": m4["B"] * world, "steps": k4, "value": m4["B"] * world * k4 / (m4["ms"] * 1e-3),
                  "unit": "scenes/s", "ms_per_step": m4["ms"] / k4,
                  "e2e": {"value": m4["B"] * world * k4 / (m4["ms_e2e"] * 1e-3), "unit": "scenes/s",
                          "ms_per_step": m4["ms_e2e"] / k4, "h2d_bytes_per_step": m4["h2d"], "d2h_bytes_per_step": 4,
                          "last_loss": m4["last"]}}
    RA.assert_clean_process()
    if rank != 0:
        finish(world)
        return
    scenes = m["B"] * world * args.steps
    line = {
        "impl": "reference", "metric": METRIC, "value": scenes / (m["ms"] * 1e-3), "unit": "scenes/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.config, m["B"], m["N"], m["F"], m["num_words"]),
                   "global_batch": m["B"] * world, "points": m["N"], "point_floats": m["F"], "tf32": False,
                   "l2": "256 MB buffer rewritten between iterations (inside the timed region)",
                   "parallelism": "dp%d" % world},
        "issue": "eager launches (the reference's own solver loop)",
        "reference": ("UNMODIFIED reference Python (baseline/_ref: models/*.py, lib/pointnet2/*.py, lib/loss_helper.py) "
                      "over the UNMODIFIED lib/pointnet2 CUDA kernels (sm_100 build, oracle/_ref); PyG's "
                      "MessagePassing.propagate restated (baseline/shims.py: the one substitution); "
                      "CUDA_LAUNCH_BLOCKING unset; N>1 = independent replicas, no collective"),
        "e2e": {"value": scenes / (m["ms_e2e"] * 1e-3), "unit": "scenes/s", "ms_per_step": m["ms_e2e"] / args.steps,
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 4, "last_loss": m["last"]},
        "gpu_launches": 0,
        "clocks": m["clocks"],
    }
    if second is not None:
        line["config4"] = second
    if m.get("predict") is not None:
        line["predict"] = m["predict"]
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the secondary BASELINE configs[3] measurement")
    ap.add_argument("--no-predict", action="store_true", help="skip the benchmark/predict.py-shaped measurement")
    ap.add_argument("--no-graph", action="store_true", help="ours: issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--ref-device", default="auto", choices=["auto", "cuda", "cpu"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

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
    if args.impl == "reference":
        run_reference(args, emit, rank, world, local_rank)
    else:
        run_ours(args, emit, rank, world, local_rank)


if __name__ == "__main__":
    main()
