#!/usr/bin/env python
"""bench.py -- depth frames/s of the ESTDepth plane-sweep + EST inference hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1], "cfg2"): one 5-frame 480x640 window, D=64 depth planes, ResNet-50 context
encoder, Joint mode, STEADY-STATE window (window 2 of a scene: EST fusion active with one memory volume,
SURVEY.md 8d) -> 3 depth maps per step.  Synthetic images / poses / random-init weights (estdepth_b200.synth).
One process per GPU; every rank runs its own independent sequence (weak scaling, no data-path collective:
sequences are independent units); value = total depth maps / max-over-ranks device time.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same step through the public call
with host buffers -- pinned H2D of the window's images/poses/intrinsics and D2H of the depth maps a driver saves
(eval_hybrid.py:259-286) inside the timed region.  `roofline`: the dominant kernel (3-D convolution) timed live
with CUDA events; `kernels`: the same for every kernel family, incl. the HBM roofline of the fused warp->cost
kernel.  `cpu_baseline`: the oracle (a port of the reference's algorithm) timed on this box's host cores on a
bounded sample.
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


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def workload_string(workload):
    V, H, W, D, resnet = WORKLOADS[workload]
    return ("%s: %d-frame %dx%d Joint window, D=%d, ResNet-%d, steady-state EST window (1 memory volume), %d depth maps/step, "
            "1 sequence per GPU" % (workload, V, H, W, D, resnet, V - 2))


def oracle_sample(workload, steps, warmup, budget_s=150.0):
    """Times the oracle (CPU port of the reference's algorithm) on the bench workload itself.

    One step = ONE steady-state window of the workload (cfg2: 5 frames -> 3 depth maps, EST fusion of every target with the
    two other targets and one memory volume), exactly what a step of the GPU arm computes; the memory volume is synthetic
    (the arithmetic does not depend on its values).  About 8 s of CPU per window on 16 cores.  (Round 1 first sampled one
    3-frame ESTM step instead: 3.9 s per depth map against 4.7 s per map for the full window on 8 cores -- a 20 % flattering
    of the baseline.)  The run stops early once ``budget_s`` is spent, with at least one timed step.
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
    V, H, W, D, resnet = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": base["steps"],
            "warmup": max(0, min(args.warmup, 1)), "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload), "arm": "CPU oracle port of the reference's algorithm, all host cores"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- this repo's arm (GPU)
def kernel_rooflines(prof, workload, peaks):
    """Per-kernel-family achieved throughput from the CUDA-event profile pass; algorithmic bytes/flops of SURVEY.md 8(d)."""
    V, H, W, D, _ = WORKLOADS[workload]
    P = (H // 4) * (W // 4)
    Vx = D * P
    out = {}
    for name, (ms, calls, flops, bytes_) in prof.items():
        if calls == 0:
            continue
        avg_s = ms / calls * 1e-3
        entry = {"calls_per_step": calls, "avg_us": avg_s * 1e6, "share_ms_per_step": ms}
        if bytes_:
            entry["algorithmic_MB"] = bytes_ / calls / 1e6
            entry["GBps"] = bytes_ / calls / avg_s / 1e9
            entry["hbm_frac"] = entry["GBps"] / peaks["hbm"]
        if flops:
            entry["algorithmic_GFLOP"] = flops / calls / 1e9
            entry["TFLOPps"] = flops / calls / avg_s / 1e12
        out[name] = entry
    return out


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
    torch.backends.cudnn.benchmark = True                  # as shipped (eval_hybrid.py:13)
    torch.backends.cudnn.allow_tf32 = False                # strict fp32 in the cuDNN feeders: parity gate is 1e-3
    torch.backends.cuda.matmul.allow_tf32 = False

    from estdepth_b200 import DepthNetHybrid, synth, ops, _lib
    V, H, W, D, resnet = WORKLOADS[args.workload]
    T = V - 2
    model = DepthNetHybrid(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet,
                           **({"precision": args.precision} if args.precision else {}),
                           **({"geometry": args.geometry} if args.geometry else {}))
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
    model.eval().to(dev)

    # window 1 (frames 0..4) primes the hidden state, window 2 (frames 3..7) is the timed steady-state step
    seed = 100 * rank
    w1 = synth.synth_inputs(V, H, W, seed=seed, start=0)
    w2 = synth.synth_inputs(V, H, W, seed=seed, start=V - 2)
    host = [t.pin_memory() for t in w2[:3]]
    dev_in = [t.to(dev, non_blocking=True) for t in host]
    # camera poses / intrinsics (5 x 4x4 + 3x3 floats) are host-side metadata in both arms: the model derives the warps'
    # matrices from them with the reference's own torch ops on the host, which is what the parity tests pin
    # (tests/run_fullsize_parity.py); the images are what "resident" refers to
    _, state, pstate = model(w1[0].to(dev), w1[1], w1[2], None, mode="val")

    def step_resident():
        return model(dev_in[0], host[1], host[2], None, state, pstate, mode="val")

    save_keys = [("depth", t, s) for t in range(T) for s in (2, 0)]        # what eval_hybrid.py writes out
    host_out = [torch.empty(1, 1, H, W).pin_memory() for _ in save_keys]

    def step_e2e():
        imgs_dev = host[0].to(dev, non_blocking=True)
        outputs, _, _ = model(imgs_dev, host[1], host[2], None, state, pstate, mode="val")
        for buf, key in zip(host_out, save_keys):
            buf.copy_(outputs[key], non_blocking=True)
        torch.cuda.current_stream().synchronize()         # the driver consumes the maps before the next window

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    # nvidia-smi is started BEFORE the warm-up and the warm-up lasts until it has had time to initialise NVML and take its first
    # samples, and until a GPU that was idle has ramped its clocks (>= W steps and >= 1 s of the same load): started right at the
    # timed region, its start-up (a process spawn + driver queries) stalled the first timed launches by several ms, and a bench
    # that was the first work on a fresh box timed its first steps on a GPU still leaving its idle power state -- together
    # 1 ms per step at K = 5 (profiles/ab_r01.txt, bench_v1 vs bench_u1).  Every sample is taken under load.
    sampler = ClockSampler(local) if rank == 0 else None
    t_warm, n_warm = time.perf_counter(), 0
    while n_warm < max(3, args.warmup) or (n_warm < 120 and time.perf_counter() - t_warm < 1.0):
        step_resident()
        n_warm += 1
        if n_warm >= max(3, args.warmup):
            torch.cuda.current_stream().synchronize()      # the second is GPU time, not enqueue time
    launches0 = _lib.launch_count()
    ms_resident = timed(step_resident, args.steps)
    launches = (_lib.launch_count() - launches0) // args.steps
    clocks = sampler.stop() if sampler else None
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # host issue time of a step (informational): two steps enqueued on an idle GPU without waiting -- few enough launches to
    # fit the driver's launch queue, so the host is not throttled by the device.  Well below ms_per_step = the GPU is the
    # bottleneck and the host runs ahead; close to it = the step is launch/host bound.
    barrier()
    import time as _time
    t0 = _time.perf_counter()
    step_resident()
    step_resident()
    host_issue_ms = (_time.perf_counter() - t0) * 1e3 / 2
    barrier()

    # profile pass: CUDA events around every kernel family of the library (same stream, same shapes)
    # (the context branch stays on the main stream for this pass: with it running beside the 3-D kernels an event pair would
    # time a kernel that shares the SMs, and the roofline wants the kernel alone)
    ops.PROFILE = ops.KernelProfile()
    overlap, model.overlap_context = model.overlap_context, 0
    barrier()
    prof_steps = max(1, min(args.steps, 5))
    for _ in range(prof_steps):
        step_resident()
    torch.cuda.synchronize()
    prof = ops.PROFILE.summary(prof_steps)
    ops.PROFILE = None
    model.overlap_context = overlap

    # isolated timing of the two kernels the roofline targets name: back-to-back launches (events bracket the whole batch,
    # so the ~3 us per-launch event/launch gap of the in-step profile is amortised); outputs rotate over buffers > L2
    isolated = {}
    if rank == 0:
        L = model._layers(dev)
        Hq, Wq = H // 4, W // 4
        g = torch.Generator(device="cpu").manual_seed(5)
        maps = [torch.randn(8, Hq, Wq, 4, generator=g).to(dev) for _ in range(2)]
        homo = ops.homography_setup(dev_in[1][0, 1].contiguous(), dev_in[1][0, 0].contiguous(),
                                    model.scale_cam_intr(dev_in[2], 0.25)[0].contiguous())
        vols = [torch.empty(8, D, Hq, Wq, 4, device=dev) for _ in range(3)]
        dvals = model._depth_dev

        def batch(fn, n=30):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e-3

        t = batch(lambda i: ops.warp_cost(maps[0], maps[1], homo, dvals, vols[i % 3]))
        nbytes = 4.0 * 32 * Hq * Wq * (D + 2)
        isolated["warp_cost"] = {"us": t * 1e6, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / t / 1e9}
        vols[0].normal_()
        t = batch(lambda i: ops.conv3d(L["dres0.0"], vols[0], vols[1 + i % 2], precision=model.precision))
        flops = 54.0 * 32 * 32 * D * Hq * Wq
        isolated["conv3d_32to32"] = {"us": t * 1e6, "algorithmic_GFLOP": flops / 1e9, "TFLOPps": flops / t / 1e12}
        del vols, maps

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    peaks = measured_peaks()
    frames = T * world * args.steps
    value = frames / (ms_resident * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)
    kernels = kernel_rooflines(prof, args.workload, peaks)
    for name, iso in isolated.items():
        if "GBps" in iso:
            iso["hbm_frac"] = iso["GBps"] / peaks["hbm"]
        target = "warp_cost" if name == "warp_cost" else "conv3d_" + model.precision
        if target in kernels:
            kernels[target]["isolated_" + name] = iso
    dom = max(kernels.items(), key=lambda kv: kv[1]["share_ms_per_step"])
    conv_name = "conv3d_" + model.precision
    conv = kernels.get(conv_name, dom[1])
    mma_per_flop = {"fp32": 0.0, "3xtf32": 3.0, "3xf16": 3.0, "3xf16r": 3.0, "3xf16r2": 3.0}[model.precision]
    # dram__bytes_read.sum + dram__bytes_write.sum of one 32->32 launch at cfg2 size from the committed ncu --set full capture
    # (profiles/kernels_r01_final.txt): CTA-pair kernel 217.5 + 115.2 MB, single-CTA ring kernel 256.9 + 119.0 MB, against
    # 314.6 MB algorithmic (157.3 MB read + 157.3 MB written; the reads carry the 18x34 / 16x32 halo, part of the output is
    # still in L2 when the kernel ends)
    traffic = {"3xf16r2": 332.6e6, "3xf16r": 375.8e6}.get(model.precision) if args.workload == "cfg2" else None
    kname = {"3xf16r": "estd::ring::conv3d_ring_kernel (3x3x3 implicit GEMM on tcgen05, plane-ring schedule, fp16 two-term split)",
             "3xf16r2": "estd::ring2::conv3d_ring2_kernel (3x3x3 implicit GEMM on tcgen05, plane-ring schedule on CTA pairs / "
                        "cta_group::2, fp16 two-term split)"}.get(
        model.precision, "estd conv3d kernel (%s)" % model.precision)
    roofline = {"kernel": kname if conv_name in kernels else dom[0], "bound": "tensor",
                "achieved": conv.get("TFLOPps"), "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": (conv.get("TFLOPps") or 0.0) / peaks["bf16_sustained"], "traffic": traffic,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "note": "achieved = ALGORITHMIC fp32 conv flops (54*Cin*Cout*Vx) / CUDA-event time, averaged over the step's launches; "
                        "the error-compensated split issues %.0fx that many tensor-core flops (tensor_flops_issued_TFLOPps / peak = the "
                        "tensor-pipe fraction); peak = dense bf16 (cuBLAS).  ncu on the CTA-pair kernel: tensor pipe 82-88 %% active; the single-"
                        "CTA ring kernel is bound by the shared-memory pipe that feeds the tensor core (an M=128 N=96 K=16 MMA needs 56 "
                        "wavefronts of operands for 48 cycles of math; the pair needs 44), profiles/README.md. share of step = %.1f%%" % (mma_per_flop, 100.0 * conv["share_ms_per_step"] / sum(k["share_ms_per_step"] for k in kernels.values())),
                "tensor_flops_issued_TFLOPps": (conv.get("TFLOPps") or 0.0) * mma_per_flop,
                "tensor_pipe_frac_issued": (conv.get("TFLOPps") or 0.0) * mma_per_flop / peaks["bf16_sustained"]}
    cpu_base, _ = (None, None)
    if not args.no_cpu_baseline and world == 1:             # reported at N = 1 only (rank 0's host cores)
        cpu_base, _ = oracle_sample(args.workload, 1, 1, budget_s=60.0)
    h2d = host[0].numel() * host[0].element_size() + 4 * (6 * 12 + 9 * 30)      # images + the warps' matrix tables
    d2h = sum(t.numel() * t.element_size() for t in host_out)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_resident / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (conv3d: %s)" % model.precision, "data": "synthetic",
            "config": {"workload": workload_string(args.workload),
                       "l2": "volumes are 157 MB each (> 126 MB L2); no explicit flush", "cudnn_tf32": False,
                       "warmup_steps_run": n_warm,
                       "camera_parameters": "host tensors (matrices of the warps derived on the host with the reference's torch ops)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue_ms, "clocks": clocks, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_base}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp32", "3xtf32", "3xf16", "3xf16r", "3xf16r2"], help="conv3d arithmetic (default: the model's)")
    ap.add_argument("--geometry", default=None, choices=["torch", "fp64"], help="camera-matrix derivation (default: the model's)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~1 min oracle timing on the host cores")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
