#!/usr/bin/env python
"""bench.py -- depth frames/s of the ESTDepth plane-sweep + EST inference hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1], "cfg2"): one 5-frame 480x640 window, D=64 depth planes, ResNet-50 context encoder,
Joint mode, STEADY-STATE window (window 2 of a scene: EST fusion active with one memory volume) -> 3 depth
maps per step.  Synthetic images / poses / random-init weights (estdepth_b200.synth).  One process per GPU; every rank runs
its own independent sequence (weak scaling: sequences are independent units); at N > 1 every step ends with the
one exchange the path has in data-parallel mode -- the NCCL all_gather of the depth maps a driver saves (BASELINE configs[3]) --
inside the timed region.  value = total depth maps / max-over-ranks device time.

Timing: after the warm-up, R = 5 repetitions of EXACTLY K steps each, every repetition bracketed by barrier +
torch.cuda.synchronize() and timed with CUDA events (max over ranks); the line reports the MEDIAN repetition
(``ms_per_step``, ``value``) and all of them (``repeats_ms_per_step``).  Resident and end-to-end repetitions alternate.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same step through the public call with host
buffers -- pinned H2D of the window's images and D2H of the depth maps a driver saves (eval_hybrid.py:259-286) inside the
timed region.  `roofline`: the dominant kernel (3-D convolution) timed live with CUDA events; `kernels`: the same for every
kernel family, incl. the HBM rooflines of the fused warp->cost kernel, the EST attention gather and the soft-argmin.
`cpu_baseline`: the oracle (a port of the reference's algorithm) timed on this box's host cores on a bounded sample.
`extras` (N = 1): the drivers' real call with CUDA camera parameters, BASELINE cfg3 (ESTM 20-frame clip), cfg5 (640x960,
D=128) and the reference algorithm as plain PyTorch-CUDA ops on the same GPU (north_star's >= 10x denominator);
(N > 1): BASELINE cfg4 (32 sequences partitioned over the ranks + gather, checked against rank 0's own run of all 32) and
the cfg3 clip cut over the ranks with the NCCL hidden-state hand-off (checked bit for bit against the sequential loop).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "depth frames/sec (5-frame seq, 480x640, D=64)"
UNIT = "frames/s"
WORKLOADS = {
    # name: (views, height, width, ndepths, resnet)
    "cfg2": (5, 480, 640, 64, 50),
    "cfg1": (5, 128, 160, 32, 18),
    "cfg5": (5, 640, 960, 128, 50),
}
REPEATS = 5


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, reasons, mx = [], set(), None
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.tmp.close()
        os.unlink(self.tmp.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload_string(workload):
    V, H, W, D, resnet = WORKLOADS[workload]
    return ("%s: %d-frame %dx%d Joint window, D=%d, ResNet-%d, steady-state EST window (1 memory volume), %d depth maps/step, "
            "1 sequence per GPU" % (workload, V, H, W, D, resnet, V - 2))


def config_dict(workload):
    """The SAME dict in both arms (the driver compares them)."""
    return {"workload": workload_string(workload),
            "inputs": "synthetic low-passed noise images, synthetic camera track, seeded random-init weights (estdepth_b200.synth)",
            "precision": "fp32 inputs / outputs / parity gate; strict fp32 (no TF32) in both arms",
            "l2": "volumes are 157 MB each (> 126 MB L2); no explicit flush",
            "camera_parameters": "host tensors (the warps' matrices are derived with the reference's torch ops on the host)"}


def median(xs):
    xs = sorted(xs)
    n = len(xs)
    return xs[n // 2] if n % 2 else 0.5 * (xs[n // 2 - 1] + xs[n // 2])


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def oracle_sample(workload, steps, warmup, budget_s=150.0):
    """Times the oracle (CPU port of the reference's algorithm) on the bench workload itself.

    One step = ONE steady-state window of the workload (cfg2: 5 frames -> 3 depth maps, EST fusion of every target with the
    two other targets and one memory volume), exactly what a step of the GPU arm computes; the memory volume is synthetic
    (the arithmetic does not depend on its values).  About 8 s of CPU per window on 16 cores.  The run stops early once
    ``budget_s`` is spent, with at least one timed step.
    """
    from estdepth_b200 import synth
    from estdepth_b200.model import DepthNetHybrid
    from oracle import estdepth_oracle as orc
    V, H, W, D, resnet = WORKLOADS[workload]
    T = V - 2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tmpl = DepthNetHybrid(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet).state_dict()
    sd = synth.synth_state_dict(tmpl, seed=0)
    cfg = dict(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet, est=True)
    imgs, poses, K, _ = synth.synth_inputs(V, H, W, seed=0, start=V - 2)
    g = torch.Generator().manual_seed(11)
    state = {"keys": [torch.relu(torch.randn(1, 16, D, H // 4, W // 4, generator=g))],
             "values": [torch.tanh(torch.randn(1, 16, D, H // 4, W // 4, generator=g))]}
    mem_pose = [synth.camera_track(1, start=V - 3)]
    times = []
    t_begin = time.perf_counter()
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.forward(sd, cfg, imgs, poses, K, state, mem_pose)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
                break
            if time.perf_counter() - t_begin > budget_s and i + 1 >= 1 and not times:
                warmup = i + 1          # out of budget during warm-up: the next step is the timed one
    mean = sum(times) / len(times)
    return dict(value=T / mean, unit=UNIT, cores=cores, kind="port", steps=len(times),
                sample="%d steady-state window(s) of the workload (%d frames -> %d depth maps, EST fusion with 1 memory volume) at "
                       "%dx%d D=%d R%d, fp32, torch CPU %d threads, mean of %d" % (len(times), V, T, H, W, D, resnet, cores, len(times))), mean


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    base, mean = oracle_sample(args.workload, args.steps, max(0, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": base["steps"],
            "warmup": max(0, min(args.warmup, 1)), "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload),
            "arm": "CPU oracle port of the reference's algorithm (oracle/estdepth_oracle.py), all host cores; the Python reference "
                   "itself cannot travel to the GPU box",
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- this repo's arm (GPU)
INCLUSIVE = "(inclusive)"


def kernel_rooflines(prof, peaks):
    """Per-kernel-family achieved throughput from the CUDA-event profile pass; algorithmic bytes/flops of SURVEY.md 8(d).
    Stage brackets appear with their EXCLUSIVE time (torch / cuDNN ops that are not one of the library's families); their
    inclusive time is listed under ``<name>(inclusive)`` and is not part of any share."""
    out = {}
    total = sum(ms for name, (ms, _, _, _) in prof.items() if not name.endswith(INCLUSIVE))
    for name, (ms, calls, flops, bytes_) in prof.items():
        if name.endswith(INCLUSIVE):
            out[name] = {"ms_per_step": ms}
            continue
        if calls == 0:
            continue
        avg_s = ms / calls * 1e-3
        entry = {"calls_per_step": calls, "avg_us": avg_s * 1e6, "share_ms_per_step": ms, "share_of_step": ms / total if total else None}
        if bytes_:
            entry["algorithmic_MB"] = bytes_ / calls / 1e6
            entry["GBps"] = bytes_ / calls / avg_s / 1e9
            entry["hbm_frac"] = entry["GBps"] / peaks["hbm"]
        if flops:
            entry["algorithmic_GFLOP"] = flops / calls / 1e9
            entry["TFLOPps"] = flops / calls / avg_s / 1e12
        out[name] = entry
    return out


def batch_time(fn, n=30, warm=3):
    """Back-to-back launches bracketed by one event pair (amortises the ~3 us event / launch gap); seconds per call."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def isolated_kernels(model, dev, H, W, D, peaks):
    """The kernels the roofline targets name, timed alone: K1 warp->cost, K2 32->32, K3 EST attention (N = 1, 2, 3), K4
    soft-argmin with the fused 1x1x1 head.  Outputs rotate over buffers larger than L2."""
    from estdepth_b200 import ops, synth
    out = {}
    L = model._layers(dev)
    Hq, Wq = H // 4, W // 4
    g = torch.Generator(device="cpu").manual_seed(5)
    maps = [torch.randn(8, Hq, Wq, 4, generator=g).to(dev) for _ in range(2)]
    poses = synth.camera_track(5).to(dev)
    K4 = model.scale_cam_intr(synth.intrinsics(H, W).unsqueeze(0), 0.25)[0].to(dev).contiguous()
    homo = ops.homography_setup(poses[1].contiguous(), poses[0].contiguous(), K4)
    vols = [torch.empty(8, D, Hq, Wq, 4, device=dev) for _ in range(3)]
    dvals = model._depth_dev
    t = batch_time(lambda i: ops.warp_cost(maps[0], maps[1], homo, dvals, vols[i % 3]))
    nbytes = 4.0 * 32 * Hq * Wq * (D + 2)
    out["warp_cost"] = {"us": t * 1e6, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / t / 1e9, "hbm_frac": nbytes / t / 1e9 / peaks["hbm"]}
    vols[0].normal_()
    t = batch_time(lambda i: ops.conv3d(L["dres0.0"], vols[0], vols[1 + i % 2], precision=model.precision))
    flops = 54.0 * 32 * 32 * D * Hq * Wq
    out["conv3d_32to32"] = {"us": t * 1e6, "algorithmic_GFLOP": flops / 1e9, "TFLOPps": flops / t / 1e12,
                            "frac_of_burst_bf16": flops / t / 1e12 / peaks["bf16"], "issued_frac_of_burst_bf16": 3 * flops / t / 1e12 / peaks["bf16"]}
    del vols, maps
    # K3: key / value volumes of 3 sources + the target key; warps between neighbouring frames of the synthetic track
    kv = [torch.randn(4, D, Hq, Wq, 4, generator=g).to(dev) for _ in range(8)]
    hs = [torch.empty(4, D, Hq, Wq, 4, device=dev) for _ in range(2)]
    tabs = ops.volume_warp_tables_torch([poses[2], poses[1], poses[3], poses[0]], 1, K4)[0].contiguous()      # [3, 30]
    for n in (1, 2, 3):
        t = batch_time(lambda i: ops.est_attend(kv[0], [kv[1 + 2 * j] for j in range(n)], [kv[2 + 2 * j] for j in range(n)], tabs[:n].contiguous(),
                                                dvals, model.depth_min, model.depth_interval, out=hs[i % 2]), n=20)
        nbytes = 4.0 * 16 * D * Hq * Wq * (2 + 2 * n)
        out["est_attend_n%d" % n] = {"us": t * 1e6, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / t / 1e9, "hbm_frac": nbytes / t / 1e9 / peaks["hbm"]}
    logits = torch.empty(D, Hq, Wq, device=dev)
    dep, prob = torch.empty(H, W, device=dev), torch.empty(H, W, device=dev)
    t = batch_time(lambda i: ops.head_softargmin(dvals, hidden=kv[i % 8], head_w=L["head0_w"], head_b=L["head0_b"], logits_out=logits,
                                                 depth_out=dep, prob_out=prob, up=4))
    nbytes = 4.0 * (16 * D * Hq * Wq + D * Hq * Wq + 2 * H * W)
    out["head_softargmin"] = {"us": t * 1e6, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / t / 1e9, "hbm_frac": nbytes / t / 1e9 / peaks["hbm"]}
    return out


def make_model(workload, dev, args):
    from estdepth_b200 import DepthNetHybrid, synth
    V, H, W, D, resnet = WORKLOADS[workload]
    model = DepthNetHybrid(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet,
                           **({"precision": args.precision} if args.precision else {}),
                           **({"geometry": args.geometry} if args.geometry else {}))
    sd = synth.synth_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    return model.eval().to(dev), sd


def estm_sequential(model, frames, n_frames, window=3, memory_size=2, frame_ids=False):
    """eval_hybrid_seq.py:169-193: one forward per new frame, a memory of the last ``memory_size`` hidden states.
    ``frame_ids``: pass the frames' indices (the optional API extension: matching features of frames shared with earlier windows
    are reused instead of recomputed)."""
    from estdepth_b200 import sharding
    memory, maps = [], []
    model._feat_cache.clear()
    for s in range(n_frames - window + 1):
        pre = sharding._flatten_memory(memory)
        out, costs, cposes = model(*frames(s), None, pre[0], pre[1], mode="val",
                                   frame_ids=list(range(s, s + window)) if frame_ids else None)
        memory.append((costs, cposes))
        if len(memory) > memory_size:
            memory.pop(0)
        maps.append(torch.cat([out[("depth", 0, 2)], out[("depth", 0, 0)]], 1))
    return torch.cat(maps)


def extras_single_gpu(model, sd, dev, args, step_ms, host, state, pstate):
    """N = 1 extras (each a few seconds): see the module docstring."""
    from estdepth_b200 import synth
    V, H, W, D, resnet = WORKLOADS[args.workload]
    T = V - 2
    ex = {}

    def timed_loop(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    # (1) the drivers' real call: camera parameters on the GPU (eval_hybrid.py: tocuda(sample)) -- the matrices of the warps are
    # then derived by ~90 tiny torch launches on a side stream
    imgs_dev, poses_dev, K_dev = host[0].to(dev), host[1].to(dev), host[2].to(dev)
    pstate_dev = [p.to(dev) for p in pstate]
    ms = timed_loop(lambda: model(imgs_dev, poses_dev, K_dev, None, state, pstate_dev, mode="val"), max(5, args.steps))
    ex["cuda_camera_parameters"] = {"ms_per_step": ms, "frames_per_s": T / ms * 1e3, "vs_host_parameters": step_ms / ms}

    # (2) BASELINE cfg3: ESTM sequential mode, 20-frame clip -> 18 forwards of 3 frames, memory 2 (eval_hybrid_seq.py:169-193)
    if args.workload == "cfg2":
        n_frames = 20
        clip = [synth.synth_inputs(3, H, W, seed=0, start=s) for s in range(n_frames - 2)]
        clip = [(c[0].to(dev), c[1], c[2]) for c in clip]
        ms = timed_loop(lambda: estm_sequential(model, lambda s: clip[s], n_frames), 3)
        ms_ids = timed_loop(lambda: estm_sequential(model, lambda s: clip[s], n_frames, frame_ids=True), 3)
        ex["cfg3_estm"] = {"clip_frames": n_frames, "forwards": n_frames - 2, "ms_per_clip": ms, "ms_per_forward": ms / (n_frames - 2),
                           "frames_per_s": (n_frames - 2) / ms * 1e3, "note": "images resident; hidden-state carry through the public forward()",
                           "with_frame_ids": {"ms_per_clip": ms_ids, "frames_per_s": (n_frames - 2) / ms_ids * 1e3,
                                              "note": "optional frame_ids= extension (SURVEY 8f rank 1): every frame's matching features computed "
                                                      "once per clip instead of up to 3 times; same depth maps"}}
        del clip

    # (3) the reference ALGORITHM as plain PyTorch-CUDA ops on this GPU (the oracle's op sequence with every tensor on the
    # device = the reference's eval path: cuDNN convolutions, ATen grid_sample, ~60 host-syncing inverses) -- north_star's
    # ">= 10x the reference PyTorch-CUDA eval frames/sec" denominator.  Baseline only.
    if not args.no_torch_cuda_baseline:
        from oracle import estdepth_oracle as orc
        sd_dev = {k: v.to(dev) for k, v in sd.items()}
        cfg = dict(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet, est=True)
        base = {}
        flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
        torch.set_default_device(dev)
        try:
            with torch.no_grad():
                for name, tf32 in (("as_shipped_cudnn_tf32", True), ("strict_fp32", False)):
                    torch.backends.cudnn.allow_tf32 = tf32
                    torch.backends.cuda.matmul.allow_tf32 = False
                    torch.backends.cudnn.benchmark = True                     # eval_hybrid.py:13
                    orc.forward(sd_dev, cfg, imgs_dev, poses_dev, K_dev, state, pstate_dev)
                    ms = timed_loop(lambda: orc.forward(sd_dev, cfg, imgs_dev, poses_dev, K_dev, state, pstate_dev), 4)
                    base[name] = {"ms_per_step": ms, "frames_per_s": T / ms * 1e3, "speedup_of_this_repo": ms / step_ms}
        finally:
            torch.set_default_device("cpu")
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = flags
        base["what"] = "oracle/estdepth_oracle.py (the reference's op sequence) with all tensors on cuda:0, same window, same weights"
        ex["torch_cuda_baseline"] = base
        del sd_dev
        torch.cuda.empty_cache()

    # (4) BASELINE cfg5: 640x960, D=128 -- frames/s and the fused warp->cost kernel against the HBM roofline
    if args.workload == "cfg2" and not args.no_cfg5:
        peaks = measured_peaks()
        m5, _ = make_model("cfg5", dev, args)
        V5, H5, W5, D5, _ = WORKLOADS["cfg5"]
        a = synth.synth_inputs(V5, H5, W5, seed=0, start=0)
        b = synth.synth_inputs(V5, H5, W5, seed=0, start=V5 - 2)
        _, st5, ps5 = m5(a[0].to(dev), a[1], a[2], None, mode="val")
        b0 = b[0].to(dev)
        ms = timed_loop(lambda: m5(b0, b[1], b[2], None, st5, ps5, mode="val"), 5)
        iso = isolated_kernels(m5, dev, H5, W5, D5, peaks)
        ex["cfg5"] = {"workload": workload_string("cfg5"), "ms_per_step": ms, "frames_per_s": (V5 - 2) / ms * 1e3,
                      "warp_cost": iso["warp_cost"], "conv3d_32to32": iso["conv3d_32to32"], "est_attend_n3": iso["est_attend_n3"]}
        m5.check()
        del m5, st5, ps5
        torch.cuda.empty_cache()
    return ex


def extras_multi_gpu(model, dev, args, rank, world):
    """N > 1 extras: BASELINE cfg4 and the cfg3 clip pipeline, both with their collective inside the timed region."""
    import torch.distributed as dist
    from estdepth_b200 import sharding, synth
    V, H, W, D, resnet = WORKLOADS[args.workload]
    T = V - 2
    ex = {}

    def synced_ms(fn):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    # ---- cfg4: 32 sequences (steady-state windows) partitioned over the ranks, depth maps all_gather'ed (sizes known from
    # the partition: one collective, no host read)
    n_seq = args.cfg4_sequences
    lo, hi = sharding.partition(n_seq, world, rank)
    counts = [b - a for a, b in (sharding.partition(n_seq, world, r) for r in range(world))]

    def prime(seq):
        w1 = synth.synth_inputs(V, H, W, seed=1000 + seq, start=0)
        w2 = synth.synth_inputs(V, H, W, seed=1000 + seq, start=V - 2)
        _, st, ps = model(w1[0].to(dev), w1[1], w1[2], None, mode="val")
        return w2[0].to(dev), w2[1], w2[2], st, ps

    def run_block(block):
        maps = []
        for imgs, poses, K, st, ps in block:
            out, _, _ = model(imgs, poses, K, None, st, ps, mode="val")
            maps.append(torch.stack([out[("depth", t, s)][0, 0] for t in range(T) for s in (2, 0)]))
        return torch.stack(maps) if maps else torch.zeros(0, 2 * T, H, W, device=dev)

    mine = [prime(s) for s in range(lo, hi)]
    run_block(mine[:1])
    ms, gathered = synced_ms(lambda: torch.cat(sharding.gather_maps(run_block(mine), counts=counts)))
    info = {"sequences": n_seq, "per_rank": counts, "ms": ms, "frames_per_s": n_seq * T / ms * 1e3,
            "gather_bytes_per_rank": int(max(counts) * 2 * T * H * W * 4), "collective": "one NCCL all_gather_into_tensor of the saved maps, inside the timed region"}
    if rank == 0:
        # the same 32 sequences on ONE rank: the gathered maps must be what a single process computes
        worst, equal = 0.0, True
        for s in range(n_seq):
            blk = mine[s - lo] if lo <= s < hi else prime(s)
            ref = run_block([blk])[0]
            worst = max(worst, float((ref - gathered[s]).abs().max()))
            equal = equal and bool(torch.equal(ref, gathered[s]))
        info["max_abs_diff_vs_single_rank"] = worst
        info["bit_identical_to_single_rank"] = equal
    ex["cfg4"] = info
    del mine, gathered
    torch.cuda.empty_cache()

    # ---- cfg3 clip pipeline: ONE 20-frame ESTM sequence cut into contiguous clips over the ranks; the last memory_size hidden
    # states (2 x 78.6 MB each) travel rank -> rank+1 over NCCL p2p as one message
    n_frames = 20

    def frames(s):
        # camera parameters on the GPU (what the drivers pass): a state received from another rank carries a CUDA pose, and
        # the comparison below needs the sequential loop to derive its matrices on the same device as the pipeline
        imgs, poses, K, _ = synth.synth_inputs(3, H, W, seed=0, start=s)
        return imgs.to(dev), poses.to(dev), K.to(dev)

    cache = {}

    def frames_cached(s):
        if s not in cache:
            cache[s] = frames(s)
        return cache[s]

    lo, hi = sharding.clip_steps(n_frames, 3, world, rank)
    for s in range(lo, hi):
        frames_cached(s)
    # every local step may be prepared ahead of the predecessor's memory (9 x 157 MB of key / value volumes at N = 2)
    pipe = sharding.EstmClipPipeline(model, window=3, memory_size=2, max_ahead=n_frames)

    def run_pipe():
        (a, b), results = pipe.run(n_frames, frames_cached, (1, 16, D, H // 4, W // 4), dev)
        local = torch.cat([torch.cat([r[("depth", 0, 2)], r[("depth", 0, 0)]], 1) for r in results]) if results else torch.zeros(0, 2, H, W, device=dev)
        steps = [q[1] - q[0] for q in (sharding.clip_steps(n_frames, 3, world, r) for r in range(world))]
        return torch.cat(sharding.gather_maps(local, counts=steps))

    run_pipe()                                             # warm-up (NCCL p2p channels)
    ms, piped = synced_ms(run_pipe)
    wait = torch.tensor([pipe.recv_wait_ms()], device=dev, dtype=torch.float64)
    dist.all_reduce(wait, op=dist.ReduceOp.MAX)
    info = {"clip_frames": n_frames, "forwards": n_frames - 2, "ms_per_clip": ms, "frames_per_s": (n_frames - 2) / ms * 1e3,
            "max_recv_wait_ms": float(wait.item()), "state_message_bytes": 2 * (2 * 16 * D * (H // 4) * (W // 4) * 4 + 64),
            "collective": "NCCL isend/irecv of the last 2 hidden states between neighbouring ranks + one all_gather of the maps"}
    if rank == 0:
        for s in range(n_frames - 2):
            frames_cached(s)
        estm_sequential(model, frames_cached, n_frames)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        seq = estm_sequential(model, frames_cached, n_frames)
        e1.record()
        torch.cuda.synchronize()
        info["sequential_ms_per_clip_one_gpu"] = e0.elapsed_time(e1)
        info["speedup_vs_one_gpu"] = e0.elapsed_time(e1) / ms
        info["max_abs_diff_vs_sequential"] = float((seq - piped).abs().max())
        info["bit_identical_to_sequential"] = bool(torch.equal(seq, piped))
    ex["estm_pipeline"] = info
    dist.barrier()
    return ex


def run_ours(args):
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    # cuDNN only runs the two stem convolutions / pools here.  benchmark stays off so that every rank (and every run) picks the
    # same algorithm: the multi-GPU extras compare maps across ranks bit for bit.  The library forces strict fp32 itself.
    torch.backends.cudnn.benchmark = False

    from estdepth_b200 import synth, ops, _lib, sharding
    V, H, W, D, resnet = WORKLOADS[args.workload]
    T = V - 2
    model, sd = make_model(args.workload, dev, args)

    # window 1 (frames 0..4) primes the hidden state, window 2 (frames 3..7) is the timed steady-state step
    seed = 100 * rank
    w1 = synth.synth_inputs(V, H, W, seed=seed, start=0)
    w2 = synth.synth_inputs(V, H, W, seed=seed, start=V - 2)
    host = [t.pin_memory() for t in w2[:3]]
    dev_in = [t.to(dev, non_blocking=True) for t in host]
    # camera poses / intrinsics (5 x 4x4 + 3x3 floats) are host-side metadata in both arms: the model derives the warps'
    # matrices from them with the reference's own torch ops on the host (extras.cuda_camera_parameters times the other way)
    _, state, pstate = model(w1[0].to(dev), w1[1], w1[2], None, mode="val")

    save_keys = [("depth", t, s) for t in range(T) for s in (2, 0)]        # what eval_hybrid.py writes out
    host_out = [torch.empty(1, 1, H, W).pin_memory() for _ in save_keys]
    gathered = torch.empty(world * len(save_keys), H, W, device=dev) if world > 1 else None

    def gather(outputs):
        """BASELINE configs[3]: NCCL gather of the depth maps -- the one collective of the data-parallel path."""
        local_maps = torch.stack([outputs[k][0, 0] for k in save_keys])
        torch.distributed.all_gather_into_tensor(gathered, local_maps)

    def step_resident():
        outputs, _, _ = model(dev_in[0], host[1], host[2], None, state, pstate, mode="val")
        if world > 1:
            gather(outputs)

    def step_e2e():
        """One window the way the reference's drivers run it: copy in, forward, copy out, wait."""
        imgs_dev = host[0].to(dev, non_blocking=True)
        outputs, _, _ = model(imgs_dev, host[1], host[2], None, state, pstate, mode="val")
        if world > 1:
            gather(outputs)
        for buf, key in zip(host_out, save_keys):
            buf.copy_(outputs[key], non_blocking=True)
        torch.cuda.current_stream().synchronize()         # the driver consumes the maps before the next window

    from estdepth_b200.io import WindowIO
    wio = WindowIO(dev)
    pipe = {"next": None, "prev": None}

    def step_e2e_pipelined():
        """The same window through the repository's own I/O helper (estdepth_b200.io.WindowIO): every step still copies its
        images from pinned host memory and reads its six maps back, but the NEXT window's upload and the PREVIOUS window's
        download overlap this window's compute; the host waits for the previous window's maps only."""
        cur = pipe["next"] if pipe["next"] is not None else wio.upload(host[0])
        pipe["next"] = wio.upload(host[0])                # the next window's images (the same synthetic window every step)
        outputs, _, _ = model(wio.ready(cur), host[1], host[2], None, state, pstate, mode="val")
        if world > 1:
            gather(outputs)
        pending = wio.download([outputs[key] for key in save_keys])
        if pipe["prev"] is not None:
            pipe["prev"].result()                         # the driver consumes the previous window's maps
            wio.release(pipe["prev"])
        pipe["prev"] = pending

    def drain_e2e_pipelined():
        if pipe["prev"] is not None:
            pipe["prev"].result()
            wio.release(pipe["prev"])
        pipe["prev"] = pipe["next"] = None

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()                                        # (pipelined arm: the last window's maps are read inside the region)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    # nvidia-smi is started BEFORE the warm-up and the warm-up lasts until it has had time to initialise NVML and take its first
    # samples, and until a GPU that was idle has ramped its clocks (>= W steps and >= 1.5 s of the same load).
    sampler = ClockSampler(local) if rank == 0 else None
    t_warm, n_warm = time.perf_counter(), 0
    if world > 1:
        # a step contains a collective: every rank must run the SAME number of warm-up steps
        n_warm = max(3, args.warmup, 100)
        for _ in range(n_warm):
            step_resident()
        torch.cuda.synchronize()
    else:
        while n_warm < max(3, args.warmup) or (n_warm < 200 and time.perf_counter() - t_warm < 1.5):
            step_resident()
            n_warm += 1
            if n_warm >= max(3, args.warmup):
                torch.cuda.current_stream().synchronize()      # the wait is GPU time, not enqueue time
    # one UNTIMED repetition of every arm: with K steps enqueued back to back the host runs ahead of the GPU and the caching
    # allocator settles on a larger working set than during the single-step warm-up (side-stream tensors are only recycled once
    # their events have completed); without this the first timed repetition of each arm paid for those cudaMallocs (18 ms/step
    # against 14.0 in the other four)
    timed(step_resident, args.steps)
    timed(step_e2e_pipelined, args.steps, after=drain_e2e_pipelined)
    timed(step_e2e, args.steps)
    launches0 = _lib.launch_count()
    rep_res, rep_e2e, rep_serial = [], [], []
    for _ in range(REPEATS):                               # resident and end-to-end repetitions alternate
        rep_res.append(timed(step_resident, args.steps) / args.steps)
        rep_e2e.append(timed(step_e2e_pipelined, args.steps, after=drain_e2e_pipelined) / args.steps)
        rep_serial.append(timed(step_e2e, args.steps) / args.steps)
    launches = (_lib.launch_count() - launches0) // (3 * REPEATS * args.steps)
    clocks = sampler.stop() if sampler else None
    model.check()                                          # fp16-range flag of everything timed above
    ms_resident, ms_e2e, ms_serial = median(rep_res), median(rep_e2e), median(rep_serial)
    # host issue time of a step (informational): two steps enqueued on an idle GPU without waiting
    barrier()
    t0 = time.perf_counter()
    step_resident()
    step_resident()
    host_issue_ms = (time.perf_counter() - t0) * 1e3 / 2
    barrier()

    # profile pass: CUDA events around every kernel family of the library (same stream, same shapes; the context branch stays
    # on the main stream for this pass so that an event pair times its kernel alone)
    ops.PROFILE = ops.KernelProfile()
    overlap, model.overlap_context = model.overlap_context, 0
    barrier()
    prof_steps = max(1, min(args.steps, 5))
    for _ in range(prof_steps):
        model(dev_in[0], host[1], host[2], None, state, pstate, mode="val")
    torch.cuda.synchronize()
    prof = ops.PROFILE.summary(prof_steps)
    ops.PROFILE = None
    model.overlap_context = overlap

    peaks = measured_peaks()
    isolated = isolated_kernels(model, dev, H, W, D, peaks) if rank == 0 else {}

    extras = {}
    if not args.no_extras:
        if world > 1:
            extras = extras_multi_gpu(model, dev, args, rank, world)
        else:
            extras = extras_single_gpu(model, sd, dev, args, ms_resident, host, state, pstate)

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    frames = T * world
    value = frames / (ms_resident * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)
    kernels = kernel_rooflines(prof, peaks)
    for name, iso in isolated.items():
        target = {"warp_cost": "warp_cost", "conv3d_32to32": "conv3d_" + model.precision, "head_softargmin": "head_softargmin"}.get(name, "est_attend")
        if target in kernels:
            kernels[target]["isolated_" + name] = iso
    conv_name = "conv3d_" + model.precision
    dom = max(((k, v) for k, v in kernels.items() if "share_ms_per_step" in v), key=lambda kv: kv[1]["share_ms_per_step"])
    conv = kernels.get(conv_name, dom[1])
    mma_per_flop = {"fp32": 0.0}.get(model.precision, 3.0)
    # dram__bytes_read.sum + dram__bytes_write.sum of one plain 32->32 launch at cfg2 size from the committed ncu --set full captures
    # (two-accumulator CTA-pair kernel, profiles/conv3d_kernels_r02.txt: 195.5 + 122.0 MB, mean of six launches; round 1,
    # profiles/kernels_r01_final.txt: single-accumulator CTA-pair kernel 217.5 + 115.2 MB, single-CTA ring 375.8 MB) against 314.6 MB
    # algorithmic (157.3 MB read + 157.3 MB written; the reads carry the 18x18 halo, part of the output is still in L2 when the
    # kernel ends)
    traffic = {"3xf16r2d": 317.5e6, "3xf16r2": 332.6e6, "3xf16r": 375.8e6}.get(model.precision) if args.workload == "cfg2" else None
    kname = {"3xf16r": "estd::ring::conv3d_ring_kernel (3x3x3 implicit GEMM on tcgen05, plane-ring schedule, fp16 two-term split)",
             "3xf16r2": "estd::ring2::conv3d_ring2_kernel (3x3x3 implicit GEMM on tcgen05, plane-ring schedule on CTA pairs / "
                        "cta_group::2, fp16 two-term split)",
             "3xf16r2d": "estd::ring2::conv3d_ring2_kernel, two accumulators per ring slot (3x3x3 implicit GEMM on tcgen05, plane-ring "
                         "schedule on CTA pairs / cta_group::2, fp16 two-term split)"}.get(model.precision, "estd conv3d kernel (%s)" % model.precision)
    achieved = conv.get("TFLOPps") or 0.0
    roofline = {"kernel": kname if conv_name in kernels else dom[0], "bound": "tensor",
                "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_sustained"], "traffic": traffic,
                "traffic_note": "bytes per plain 32->32 launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, "
                                "profiles/conv3d_kernels_r02.txt); algorithmic: 314.6e6",
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "share_of_step": conv.get("share_of_step"),
                "note": "achieved = ALGORITHMIC fp32 conv flops (54*Cin*Cout*Vx) / CUDA-event time, averaged over the step's launches; the "
                        "error-compensated split issues %.0fx that many tensor-core flops (tensor_pipe_frac_issued); peak = dense bf16 "
                        "(cuBLAS).  share_of_step = this family's time / sum of all families' exclusive times (profile pass)." % mma_per_flop,
                "tensor_flops_issued_TFLOPps": achieved * mma_per_flop,
                "tensor_pipe_frac_issued": achieved * mma_per_flop / peaks["bf16_sustained"]}
    cpu_base = None
    if not args.no_cpu_baseline and world == 1:             # reported at N = 1 only (rank 0's host cores)
        cpu_base, _ = oracle_sample(args.workload, 1, 1, budget_s=60.0)
    h2d = host[0].numel() * host[0].element_size() + 4 * (6 * 12 + 9 * 30)      # images + the warps' matrix tables
    d2h = sum(t.numel() * t.element_size() for t in host_out)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (conv3d: %s)" % model.precision, "data": "synthetic",
            "config": config_dict(args.workload),
            "timing": {"repeats": REPEATS, "statistic": "median of %d repetitions of exactly %d steps, each bracketed by barrier + synchronize, "
                                                        "CUDA events, max over ranks" % (REPEATS, args.steps),
                       "repeats_ms_per_step": rep_res, "min_ms_per_step": min(rep_res), "max_ms_per_step": max(rep_res),
                       "timed_region_s": sum(rep_res) * args.steps * 1e-3, "warmup_steps_run": n_warm + args.steps,
                       "collective_in_step": "all_gather_into_tensor of the saved depth maps (%d B per rank)" % d2h if world > 1 else None},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                    "repeats_ms_per_step": rep_e2e,
                    "how": "public call model(...) with the window's images uploaded from pinned host memory and the six saved maps read back "
                           "into pinned host memory EVERY step, through estdepth_b200.io.WindowIO: the next window's upload and the previous "
                           "window's download overlap the current window's compute (the host waits for the previous window's maps); the "
                           "first upload and the last download are inside the timed region",
                    "serial": {"value": frames / (ms_serial * 1e-3), "ms_per_step": ms_serial, "repeats_ms_per_step": rep_serial,
                               "how": "the same copies issued the way the reference's drivers do: upload, forward, download, wait -- nothing overlapped"}},
            "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue_ms, "clocks": clocks, "roofline": roofline,
            "kernels": kernels, "extras": extras, "cpu_baseline": cpu_base}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp32", "3xtf32", "3xf16", "3xf16r", "3xf16r2", "3xf16r2d"], help="conv3d arithmetic (default: the model's)")
    ap.add_argument("--geometry", default=None, choices=["auto", "torch", "fp64"], help="camera-matrix derivation (default: the model's)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~1 min oracle timing on the host cores")
    ap.add_argument("--no-extras", action="store_true", help="skip the extras (cfg3 / cfg4 / cfg5 / PyTorch-CUDA baseline / clip pipeline)")
    ap.add_argument("--no-torch-cuda-baseline", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--cfg4-sequences", type=int, default=32)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
