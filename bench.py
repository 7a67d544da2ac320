#!/usr/bin/env python
"""bench.py -- scenes/sec of the RfD-Net point-cloud hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                   (the reference algorithm on the host cores: CPU oracle)

A step = one pass of the hot path over one batch of synthetic scenes per GPU:
  80k-point ScanNet-like clouds -> backbone (4 SA + 2 FP) -> voting -> 256 proposals -> ONet decoder on the dense
  32^3 lattice for all 256 proposals.  Weak scaling: every rank processes `--scenes` scenes per step, no data-path
  collective (inference shards by scene).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "scenes/sec (80k pts, 256 proposals, 32^3 occ queries)"
WORKLOAD = "full hot path: 80k-pt scene -> backbone (4 SA + 2 FP) + vote + 256 proposals -> ONet decoder 256 x 32^3"
FLOP_PER_POINT = 1312768.0


_T0 = time.perf_counter()


def _mark(what):
    """wall-clock trace of the bench's sections on stderr (the JSON line on stdout stays alone)"""
    sys.stderr.write(f"[bench +{time.perf_counter() - _T0:7.1f}s] {what}\n")
    sys.stderr.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (pinned host buffers then live on the
    NUMA node the GPU's PCIe root hangs off; measured: D2H of the logits is ~8x slower from the far node)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1]
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception as e:  # affinity is an optimisation, never a requirement
        sys.stderr.write(f"[bench] NUMA binding skipped: {e}\n")
    return 0


def build_model(device, seed=0, backbone_precision="x3", head_precision="x3", decoder_precision="fp16", graph=False):
    from rfdnet_b200.pipeline import SceneHotPath
    from rfdnet_b200.synth import seeded_fill
    net = SceneHotPath(precision=decoder_precision, backbone_precision=backbone_precision,
                       head_precision=head_precision, graph_detection=graph).eval()
    seeded_fill(net, seed)
    return net.to(device)


def make_inputs(scenes, seed0, npts=80000):
    from rfdnet_b200.synth import scannet_like_batch
    pc = torch.from_numpy(scannet_like_batch(scenes, npts, seed0=seed0))
    g = torch.Generator().manual_seed(seed0 + 7)
    codes = torch.randn(scenes * 256, 512, generator=g)
    return pc, codes


# --------------------------------------------------------------------------------------------- CPU (reference arm)
def cpu_sample(state, budget_s=30.0):
    """One bounded sample of the reference algorithm on the host cores: the full detection path on ONE 80k scene
    (C oracle with OpenMP for the index kernels, PyTorch CPU for the MLPs) + the ONet decoder, one object x 32^3
    points per call exactly like generator.py:131-141, on as many of the scene's 256 objects as fit in `budget_s`
    seconds (all 256 when they fit: then nothing is extrapolated).
    -> dict(per_scene_s, t_detect_s, t_decode_per_object_s, objects_measured, extrapolated, wall_s)"""
    from oracle import model_ref
    sd, pc, codes, grid = state["sd"], state["pc_cpu"], state["codes_cpu"], state["grid_cpu"]
    t0 = time.perf_counter()
    with torch.no_grad():
        ep = model_ref.backbone(pc[:1], sd, prefix="detection.backbone", recip=False)
        vx, vf = model_ref.voting(ep["fp2_xyz"], ep["fp2_features"], sd, prefix="detection.voting")
        model_ref.proposal(vx, vf, sd, prefix="detection.detection", recip=False)
    t1 = time.perf_counter()
    dsd = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    n = 0
    with torch.no_grad():
        while n < 256:
            model_ref.decoder(grid.unsqueeze(0), torch.zeros(1, 32), codes[n:n + 1], dsd)
            n += 1
            if time.perf_counter() - t0 > budget_s and n < 256:
                break
    t2 = time.perf_counter()
    t_det, t_dec = t1 - t0, (t2 - t1) / n
    return {"per_scene_s": t_det + 256.0 * t_dec, "t_detect_s": t_det, "t_decode_per_object_s": t_dec,
            "objects_measured": n, "extrapolated": n < 256, "wall_s": t2 - t0}


def cpu_sample_text(r):
    return (f"detection path on 1 scene of 80k pts ({r['t_detect_s']:.2f}s: C oracle + OpenMP index kernels, torch CPU "
            f"MLPs) + ONet decoder on {r['objects_measured']} of the scene's 256 objects x 32^3 pts "
            f"({r['t_decode_per_object_s']:.3f}s/object, torch CPU fp32)"
            + ("; remaining objects extrapolated at the measured per-object time" if r["extrapolated"] else
               "; the whole scene was measured, nothing extrapolated"))


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle
    from oracle import model_ref
    torch.set_num_threads(os.cpu_count() or 1)
    net = build_model("cpu")
    sd = {k: v for k, v in net.state_dict().items()}
    pc, codes = make_inputs(1, 0)
    state = {"sd": sd, "pc_cpu": pc, "codes_cpu": codes, "grid_cpu": model_ref.make_3d_grid(32, 1.1)}
    # every step is a bounded sample; the whole run must end within a few minutes
    budget = max(4.0, min(40.0, 240.0 / (args.steps + 1)))
    cpu_sample(state, budget_s=min(budget, 8.0))  # one short warm-up sample (thread pools, allocator)
    res = [cpu_sample(state, budget_s=budget) for _ in range(args.steps)]
    per_scene = float(np.mean([r["per_scene_s"] for r in res]))
    cores = oracle.num_threads()
    best = max(res, key=lambda r: r["objects_measured"])
    val = 1.0 / per_scene
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "scenes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_scene * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scenes_per_gpu_per_step": 1, "points": 80000, "proposals": 256, "grid": 32},
            "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": cores, "kind": "port",
                             "sample": "per step: " + cpu_sample_text(best),
                             "objects_measured": int(np.min([r["objects_measured"] for r in res])),
                             "extrapolated": bool(any(r["extrapolated"] for r in res)),
                             "measured_wall_ms_per_step": float(np.mean([r["wall_s"] for r in res])) * 1e3,
                             "warmup_done": 1},
            "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank, world, local):
    from rfdnet_b200 import _lib, dist as D
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(visible.split(",")[local]) if visible and visible.split(",")[local].isdigit() else local
    bind_to_gpu_numa_node(phys)
    S = args.scenes
    net = build_model(dev, backbone_precision=args.backbone_precision, head_precision=args.head_precision,
                      decoder_precision=args.decoder_precision, graph=args.graph_detection)
    sets = [make_inputs(S, 1000 * rank + 100 * i) for i in range(2)]  # two rotating input sets
    dev_sets = [(pc.to(dev), codes.to(dev)) for pc, codes in sets]
    host_sets = [(pc.pin_memory(), codes.pin_memory()) for pc, codes in sets]
    logits_host = torch.empty((S * 256, 32768), dtype=torch.float32).pin_memory()
    hbm, tc_burst, tc_sust, peak_src = peaks()

    def step_dev(i):
        pc, codes = dev_sets[i & 1]
        return net(pc, codes)

    _mark("model + inputs built")
    # ---- warm-up
    for i in range(args.warmup):
        step_dev(i)
    torch.cuda.synchronize()

    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.TIMERS = []
    D.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_dev(i)
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    timers, _lib.TIMERS = _lib.TIMERS, None
    clocks = sampler.stop() if rank == 0 else None
    ms_max = D.max_over_ranks(ms, dev)

    _mark("device-resident timed region done")
    # ---- timed region 2: end to end through the public API with host buffers.  The step's result is what
    # Generator3D returns to its caller -- the meshes (generator.py:145-168), extracted on the device -- plus the proposal
    # scores; the variant that ships every logit to the host (round 1's e2e) is timed as well.
    def e2e_run(result):
        for i in range(3):
            net.run_host(*host_sets[i & 1], logits_host, dev, result=result, chunks=args.e2e_chunks)
        torch.cuda.synchronize()
        D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for i in range(args.steps):
            h2d, d2h = net.run_host(*host_sets[i & 1], logits_host, dev, result=result, chunks=args.e2e_chunks)
            torch.cuda.synchronize()  # the step's result is on the host before the next step starts
        t = time.perf_counter() - t0
        D.barrier()
        return D.max_over_ranks(t, dev), h2d, d2h

    t_e2e, h2d, d2h = e2e_run("mesh")
    mesh_stats = {"vertices_per_step": int(net.last_meshes[0].shape[0]), "triangles_per_step": int(net.last_meshes[1].shape[0])}
    t_e2e_lg, h2d_lg, d2h_lg = e2e_run("logits")
    t_e2e_b, h2d_b, d2h_b = e2e_run("bits")

    _mark("end-to-end variants done")
    # ---- BASELINE config 5: training step with the NCCL gradient all-reduce (every rank takes part)
    train = None
    if not args.no_train:
        try:
            train = train_bench(args, rank, world, dev)
        except Exception as e:  # the inference line must survive a failure of the training block
            import traceback
            traceback.print_exc(file=sys.stderr)
            train = {"error": repr(e)[:300]}
            if world > 1:
                raise

    if rank != 0:
        return None
    # ---- per-kernel device times (CUDA events recorded around the launches inside the timed steps)
    dec = [(s.elapsed_time(e), w) for n, s, e, w in timers if n == "onet_decode"]
    dec_ms = float(np.mean([t for t, _ in dec]))
    dec_tflops = float(np.mean([w for _, w in dec])) / (dec_ms * 1e-3) / 1e12
    chain = [(s.elapsed_time(e), w) for n, s, e, w in timers if n == "mlp_chain_tc"]
    chain_tflops = (sum(w for _, w in chain) / (sum(t for t, _ in chain) * 1e-3) / 1e12) if chain else None
    fps = [s.elapsed_time(e) for n, s, e, w in timers if n == "fps"]
    value = world * S * args.steps / (ms_max * 1e-3)
    e2e = world * S * args.steps / t_e2e
    _mark("train block done")
    qg = ballquery_group_bench(net, dev_sets[0][0], hbm)
    _mark("ball query + group operator bench done")
    skip = None
    try:
        with torch.no_grad():
            ep0, _ = net.detection(dev_sets[0][0][:1].contiguous())
        skip = skip_propagation_bench(dev, dev_sets[0][0], ep0)
    except Exception as e:  # an (f)-row extra must never take the headline line down
        import traceback
        traceback.print_exc(file=sys.stderr)
        skip = {"error": repr(e)[:300]}
    _mark("SkipPropagation block done")
    fullgen = None
    try:
        fullgen = full_generation_bench(dev, dev_sets[0][0])
    except Exception as e:
        import traceback
        traceback.print_exc(file=sys.stderr)
        fullgen = {"error": repr(e)[:300]}

    _mark("full generation block done")
    # ---- CPU baseline (bounded sample, rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, range(os.cpu_count()))  # the CPU arm may use every host core
        except Exception:
            pass
        import oracle
        from oracle import model_ref
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.cpu() for k, v in net.state_dict().items()}
        state = {"sd": sd, "pc_cpu": sets[0][0], "codes_cpu": sets[0][1], "grid_cpu": model_ref.make_3d_grid(32, 1.1)}
        r = cpu_sample(state, budget_s=25.0)
        cpu = {"value": 1.0 / r["per_scene_s"], "unit": "scenes/s", "cores": oracle.num_threads(), "kind": "port",
               "sample": cpu_sample_text(r), "objects_measured": r["objects_measured"],
               "extrapolated": r["extrapolated"], "measured_wall_s": r["wall_s"]}

    _mark("cpu baseline done")
    traffic = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "onet_decode_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    dec_mode = args.decoder_precision
    line = {
        "metric": METRIC, "value": value, "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp16": "fp16", "bf16": "bf16", "fp16x3": "fp16x3"}[dec_mode], "data": "synthetic",
        "config": {"workload": WORKLOAD + " [decoder: %s tcgen05, fp32 accumulate/residual; backbone MLPs: %s; "
                               "vote/proposal MLPs: %s]" % (dec_mode, args.backbone_precision, args.head_precision),
                   "scenes_per_gpu_per_step": S, "points": 80000, "proposals": 256, "grid": 32,
                   "parallelism": f"dp{world} (scenes sharded, no collective)",
                   "detection_launch": "one CUDA graph" if args.graph_detection else "eager (45 launches)", "e2e_chunks": args.e2e_chunks,
                   "l2": "per-step working set (logits %d MB + clouds) exceeds the 126 MB L2; inputs rotate over 2 sets"
                         % (S * 256 * 32768 * 4 // 2 ** 20)},
        "roofline": {"kernel": "onet_decode_kernel", "bound": "tensor", "achieved": dec_tflops, "peak": tc_sust,
                     "unit": "TFLOP/s", "frac": dec_tflops / tc_sust, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src})", "ms_per_launch": dec_ms,
                     "flop_per_launch": float(np.mean([w for _, w in dec])),
                     "note": "algorithmic FLOP (1,312,768 per query point); fp16x3 issues 3 MMAs per algorithmic one"},
        "ballquery_group": qg,
        "mlp_chain_tc": {"achieved": chain_tflops, "unit": "TFLOP/s", "launches_per_step": len(chain) // args.steps,
                         "ms_per_step": sum(t for t, _ in chain) / args.steps if chain else None,
                         "layers": "SA1-4 (gather-fused), FP1-2, voting, vote-aggregation SA, proposal head",
                         "note": "algorithmic FLOP; mode x3 issues 3 MMAs per algorithmic one"},
        "fps": {"ms_per_step": sum(fps) / args.steps if fps else None, "launches_per_step": len(fps) // args.steps},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": t_e2e / args.steps * 1e3,
                "result": "meshes (marching cubes on the device, rfd_extract_mesh: f32 vertices + i32 triangles + "
                          "per-object ranges) + objectness scores", **mesh_stats},
        "e2e_all_logits": {"value": world * S * args.steps / t_e2e_lg, "unit": "scenes/s", "h2d_bytes_per_step": h2d_lg,
                           "d2h_bytes_per_step": d2h_lg, "ms_per_step": t_e2e_lg / args.steps * 1e3,
                           "result": "every logit (S*256 x 32^3 f32) + objectness scores"},
        "e2e_occupancy_bits": {"value": world * S * args.steps / t_e2e_b, "unit": "scenes/s", "h2d_bytes_per_step": h2d_b,
                               "d2h_bytes_per_step": d2h_b, "ms_per_step": t_e2e_b / args.steps * 1e3,
                               "result": "occupancy bit masks (1 bit per lattice point) + objectness scores"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "train": train,
        "skip_propagation": skip,
        "full_generation": fullgen,
    }
    return line


def train_bench(args, rank, world, dev):
    """BASELINE config 5: the joint ISCNet training step (detection + SkipPropagation + ONet encoder/decoder in train
    mode, Adam) on `--train-batch` synthetic 80k-point scenes PER GPU, gradients exchanged by the bucketed NCCL all-reduce
    launched from autograd hooks while backward is still running (rfdnet_b200/train.py).  All times are CUDA-event
    times, max over ranks.
      ms_per_step        whole step (zero, forward, backward + overlapped all-reduce, optimizer)
      allreduce_us       the same buckets all-reduced back to back on an otherwise idle GPU (no overlap possible)
      bus_gbs            2 (N-1)/N x bytes / allreduce_us  (ring/NVLS bus bandwidth convention)
      exposed_us         in-step time between the last backward kernel and the completion of the last bucket
      overlap_frac       1 - exposed / allreduce  (share of the exchange hidden under backward)"""
    from rfdnet_b200 import dist as D, train as T
    from rfdnet_b200.synth import scannet_like_batch, seeded_fill
    B, K, Tpts = args.train_batch, 10, 2048
    torch.manual_seed(1234 + rank)
    model = T.JointTrainStep(boxes_per_scene=K)
    seeded_fill(model, 11)
    model = model.to(dev).train()
    pc = torch.from_numpy(scannet_like_batch(B, 80000, seed0=5000 + 100 * rank)).to(dev)
    lab = T.synthetic_labels(B, 80000, K, Tpts, dev, seed=rank)
    trainer = T.Trainer(model, bucket_bytes=args.bucket_mb << 20)
    losses = []
    for _ in range(2):
        loss, _ = trainer.step(pc, lab)
        losses.append(float(loss))
    torch.cuda.synchronize()
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    exposed = []
    e0.record()
    for _ in range(args.train_steps):
        loss, _ = trainer.step(pc, lab, record=True)
        exposed.append(trainer.ev)
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    losses.append(float(loss))
    ms = D.max_over_ranks(e0.elapsed_time(e1) / args.train_steps, dev)
    exposed_us = D.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in exposed])) * 1e3, dev)
    nbytes = trainer.buckets.nbytes
    ar_us = None
    if world > 1:
        for _ in range(2):
            trainer.buckets.allreduce_only()
        torch.cuda.synchronize()
        D.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            trainer.buckets.allreduce_only()
        a1.record()
        torch.cuda.synchronize()
        ar_us = D.max_over_ranks(a0.elapsed_time(a1) / 5 * 1e3, dev)
    finite = all(np.isfinite(losses))
    out = {"workload": "joint ISCNet train step (BASELINE config 5): detection (train-mode BN) + SkipPropagation "
                       "(STN_Group r=1.0/nsample=1024 + PointSeg + ResnetPointnet) + ONet Encoder_Latent/batch-stat CBN "
                       "decoder (KL + BCE) + Adam; point-cloud ops and their scatter-add grads = librfdnet_b200, dense "
                       "layers = PyTorch library GEMMs; surrogate detection loss (models/loss.py is out of scope)",
           "scenes_per_gpu": B, "points": 80000, "boxes_per_scene": K, "occ_points_per_box": Tpts,
           "ms_per_step": ms, "scenes_per_s": world * B / (ms * 1e-3), "steps": args.train_steps, "warmup": 2,
           "params": sum(p.numel() for p in model.parameters()), "allreduce_bytes": nbytes,
           "buckets": len(trainer.buckets.buckets), "bucket_mb": args.bucket_mb,
           "allreduce_us": ar_us, "bus_gbs": (2.0 * (world - 1) / world * nbytes / (ar_us * 1e-6) / 1e9) if ar_us else None,
           "nvlink_peak_gbs": 900.0, "exposed_us": exposed_us if world > 1 else 0.0,
           "overlap_frac": (max(0.0, 1.0 - exposed_us / ar_us) if ar_us else None),
           "loss_first_last": [losses[0], losses[-1]], "loss_finite": bool(finite),
           "collective": "nccl all_reduce(AVG) per bucket, async from post-accumulate-grad hooks" if world > 1 else "none (1 GPU)"}
    del trainer, model
    torch.cuda.empty_cache()
    return out


def skip_propagation_bench(dev, pc, ep):
    """SkipPropagation.generate for the 256 proposals of ONE 80k-point scene -- STN_Group (ball query
    r = 1.0 / nsample = 1024 + heading rotation + STN3d alignment) on this library's kernels (4 launches), PointSeg and
    ResnetPointnet (hidden 512) on PyTorch's library GEMMs -- producing the 512-d shape codes the decoder consumes."""
    from rfdnet_b200 import completion
    from rfdnet_b200.synth import seeded_fill
    sp = completion.SkipPropagation(input_feature_dim=1, c_dim=512, hidden_dim=512).eval()
    seeded_fill(sp, 17)
    sp = sp.to(dev)
    pc1 = pc[:1].contiguous()
    box_xyz = ep["center"][:1].contiguous()
    heading = torch.argmax(ep["heading_scores"][:1], -1).float() * (2 * np.pi / 12)
    box_feat = torch.randn(1, 128, 256, device=dev)
    xyz = pc1[..., :3].contiguous()
    feats = torch.cat([pc1[..., 3:].transpose(1, 2), torch.zeros(1, 1, pc1.shape[1], device=dev)], 1).contiguous()

    def timed(fn, n=3):
        with torch.no_grad():
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    stn_ms, _ = timed(lambda: sp.stn(xyz, feats, box_xyz, heading))
    tot_ms, codes = timed(lambda: sp.generate(box_xyz, heading, box_feat, pc1), n=2)
    sp.fast_precision = 'fp16'
    fp16_ms, codes_h = timed(lambda: sp.generate(box_xyz, heading, box_feat, pc1), n=2)
    sp.fast_precision = None
    torch_ms, codes_t = timed(lambda: sp.generate(box_xyz, heading, box_feat, pc1), n=1)
    out = {"proposals": 256, "points": int(pc1.shape[1]), "nsample": 1024, "stn_group_ms": stn_ms, "generate_ms": tot_ms,
           "generate_fp16_operands_ms": fp16_ms, "generate_torch_layers_ms": torch_ms,
           "max_abs_diff_vs_torch_layers": float((codes - codes_t).abs().max()),
           "max_abs_diff_fp16_vs_torch_layers": float((codes_h - codes_t).abs().max()),
           "codes_shape": list(codes.shape), "codes_finite": bool(torch.isfinite(codes).all()),
           "note": "generate_ms: STN_Group (rfd_query_and_group_rotated + 2 x rfd_mlp_chain + rfd_stn_apply) + PointSeg + "
                   "ResnetPointnet on the tcgen05 chain kernel (rfd_mlp_chain_ex, x3 = fp32-grade operands, 8.6 MFLOP per "
                   "point after folding the repeated global features into per-cloud biases); generate_torch_layers_ms: the "
                   "same with PointSeg / ResnetPointnet on torch fp32 library GEMMs (the reference's formulation, 15.7 MFLOP "
                   "per point)"}
    del sp
    torch.cuda.empty_cache()
    return out


def full_generation_bench(dev, pc, scenes=2):
    """ISCNet.generate's device part (network.py:56-153) for all 256 proposals of `scenes` scenes: detection ->
    SkipPropagation (shape codes) -> ONet decoder on 32^3 -> meshes, everything on this library (pipeline.SceneGeneration).
    NOT the headline metric (BASELINE's path takes the codes as given): it shows what the widened path costs."""
    from rfdnet_b200.pipeline import SceneGeneration
    from rfdnet_b200.synth import seeded_fill
    net = SceneGeneration().eval()
    seeded_fill(net, 29)
    net = net.to(dev)
    x = pc[:scenes].contiguous()
    out = net(x, meshes=False)
    with torch.no_grad():   # centre the seeded decoder's logits so that the surfaces are not empty
        net.completion.decoder.fc_out.bias -= out["logits"].median()
    for _ in range(2):
        out = net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = net(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    v, t, r = out["meshes"].to_host()
    res = {"scenes": scenes, "proposals_per_scene": 256, "ms_per_scene": ms / scenes, "scenes_per_s": scenes / (ms * 1e-3),
           "vertices": int(len(v)), "triangles": int(len(t)),
           "stages": "detection + SkipPropagation.generate (STN_Group, PointSeg, ResnetPointnet on tcgen05) + decoder fp16 + "
                     "rfd_extract_mesh; proposal selection (NMS) is the caller's"}
    del net, out
    torch.cuda.empty_cache()
    return res


def graph_time(fn, iters, reps=3):
    """Device time of one call of `fn` (ms): `iters` calls are captured into ONE CUDA graph (after an eager warm-up on the
    capture stream, which also sizes the library's persistent workspace) and the replay is timed with CUDA events, so the
    figure is the kernels' own time back to back, not Python / launch overhead."""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    keep = []  # outputs stay alive: every captured call writes its own buffer (no L2-resident reuse of one block)
    with torch.cuda.graph(g, stream=st):
        for _ in range(iters):
            keep.append(fn())
    best = None
    with torch.cuda.stream(st):
        g.replay()
        st.synchronize()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            st.synchronize()
            t = e0.elapsed_time(e1) / iters
            best = t if best is None else min(best, t)
    torch.cuda.current_stream().wait_stream(st)
    del keep, g
    return best


def ballquery_group_bench(net, pc, hbm_peak, iters=20):
    """rfd_query_and_group (the drop-in QueryAndGroup operator) on the five layer shapes of THIS step's
    scenes: algorithmic bytes (12N + 12M + 4CN + 4(3+C)MS per scene) / CUDA-event time of `iters`
    back-to-back launches per layer, after warm-up.  The fused inference path never materialises the grouped tensor
    (mlp_chain_tc gathers through the ball-query indices), so the operator is timed here on its own."""
    from rfdnet_b200 import pointnet2_utils as pu
    bb = net.detection.backbone
    with torch.no_grad():
        ep, _ = net.detection(pc)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    agg = net.detection.detection.vote_aggregation
    shapes = [("SA1", xyz, ep["sa1_xyz"], feats, bb.sa1), ("SA2", ep["sa1_xyz"], ep["sa2_xyz"], ep["sa1_features"], bb.sa2),
              ("SA3", ep["sa2_xyz"], ep["sa3_xyz"], ep["sa2_features"], bb.sa3),
              ("SA4", ep["sa3_xyz"], ep["sa4_xyz"], ep["sa3_features"], bb.sa4),
              ("vote-agg", ep["vote_xyz"], ep["aggregated_vote_xyz"], ep["vote_features"], agg)]
    per, tot_b, tot_ms = {}, 0.0, 0.0
    for name, src, q, f, mod in shapes:
        src, q, f = src.contiguous(), q.contiguous(), f.contiguous()
        B, N, _ = src.shape
        M, C, Sn = q.shape[1], f.shape[1], mod.nsample
        ms = graph_time(lambda: pu.fused_query_and_group(src, q, f, mod.radius, Sn, True, True), iters)
        nbytes = B * (12 * N + 12 * M + 4 * C * N + 4 * (3 + C) * M * Sn)
        per[name] = {"us": ms * 1e3, "MB": nbytes / 1e6, "GB/s": nbytes / (ms * 1e-3) / 1e9,
                     "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak}
        # the product path of the same layer: ball query (indices only) + ONE tcgen05 kernel that gathers through the
        # indices, runs the shared MLP and max-pools -- the (B,3+C,M,S) grouped tensor is never written or read
        inds = torch.arange(M, dtype=torch.int32, device=src.device).expand(B, M).contiguous()
        with torch.no_grad():
            fus = graph_time(lambda: mod._forward_fused(src, f, inds, q), 10)
        per[name]["product_path"] = {"us_ball_query_plus_gather_mlp_max": fus * 1e3,
                                     "grouped_tensor_MB_not_materialised": B * 4 * (3 + C) * M * Sn / 1e6}
        tot_b += nbytes
        tot_ms += ms
    gbs = tot_b / (tot_ms * 1e-3) / 1e9
    return {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "bound": "hbm",
            "us_all_layers": tot_ms * 1e3, "per_layer": per,
            "how": f"{iters} back-to-back calls per layer captured in one CUDA graph, replay timed with CUDA events (best "
                   "of 3), incl. the grid build (SA1) and the feature transposition pass; every call writes a fresh "
                   "output (the 20 outputs of a replay exceed the 126 MB L2)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=4, help="scenes per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the config-5 training-step block")
    ap.add_argument("--graph-detection", action="store_true", help="replay the detection pass as one CUDA graph")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="object chunks of the decoder in the end-to-end call")
    ap.add_argument("--train-batch", type=int, default=8, help="scenes per GPU per training step (config 5: 8)")
    ap.add_argument("--train-steps", type=int, default=3)
    ap.add_argument("--bucket-mb", type=int, default=8, help="gradient all-reduce bucket size")
    ap.add_argument("--backbone-precision", default="x3", choices=["x3", "fp16", "bf16", "cuda"],
                    help="MLPs of SA1-4 / FP1-2 (BASELINE config 2 is fp32): x3 = split-fp16 tcgen05, fp32-grade (default); "
                         "fp16 / bf16 = single-MMA tcgen05; cuda = fp32 CUDA-core layer kernel")
    ap.add_argument("--head-precision", default="x3", choices=["x3", "fp16", "bf16", "cuda"],
                    help="voting MLP, vote-aggregation SA layer, proposal head (BASELINE config 3)")
    ap.add_argument("--decoder-precision", default="fp16", choices=["fp16", "fp16x3", "bf16"],
                    help="ONet decoder tcgen05 mode: fp16 (<= 1e-3, BASELINE config 4), fp16x3 (<= 1e-4), bf16 (legacy)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rfdnet_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "rfdnet_b200", "librfdnet_b200.so")):
        g.build()
    from rfdnet_b200 import dist as D
    # NCCL (NCCL_DEBUG=INFO) logs on STDOUT whenever it initialises something -- the communicator, NVLS on the first
    # all-reduce, the teardown.  stdout carries exactly ONE line, the JSON: fd 1 is routed to stderr for the whole run and
    # the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    line = None
    try:
        D.init_from_env("nccl")
        if world > 1:
            torch.cuda.set_device(local)
            D.barrier()
            torch.cuda.synchronize()
        line = run_gpu(args, rank, world, local)
        if world > 1:
            torch.distributed.destroy_process_group()
    finally:
        sys.stdout.flush()
        if line is not None:
            os.write(saved_stdout, (json.dumps(line) + "\n").encode())
        os.close(saved_stdout)


if __name__ == "__main__":
    main()
