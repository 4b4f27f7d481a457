"""Parity AT THE BENCHMARKED SIZES against outputs of the unmodified reference (tests/golden/*480x640*, *640x960*, made by
``python -m oracle.make_golden --fullsize`` in the build container; inputs and weights are regenerated from the same seeds).

  cfg2  5 x 480 x 640, D=64, ResNet-50: both Joint windows (no-EST then EST), head gain 3 (synth default) and 10
        (logit sigma ~ 3 -- depth error scales with the logit gain)
  cfg3  ESTM protocol at 480 x 640 (eval_hybrid_seq.py:169-193), first 4 steps
  cfg5  5 x 640 x 960, D=128: both Joint windows

plus the path the real drivers take: camera parameters as CUDA tensors, and cuDNN's TF32 switch left at PyTorch's default.
Gates (BASELINE.json north_star): every depth map within 1e-3 abs; probabilities 2e-4; hidden state 2e-4.
The worst error per output kind is printed for every configuration.
"""
import os

import numpy as np
import pytest
import torch

from estdepth_b200 import DepthNetHybrid, ops, sharding, synth
from estdepth_b200.io import DepthMapWriter
from oracle import estdepth_oracle as orc
from oracle.make_golden import FULL_STATE_STRIDE, subsample
from tests.helpers import cfg_of, synth_model_and_state

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEPTH_TOL, PROB_TOL, STATE_TOL = 1e-3, 2e-4, 2e-4


def _model(resnet, ndepths, head_gain=synth.HEAD_GAIN, **kw):
    m = DepthNetHybrid(ndepths=ndepths, depth_min=0.1, depth_max=10.0, resnet=resnet, **kw)
    sd = synth.synth_state_dict(m.state_dict(), seed=0, head_gain=head_gain)
    m.load_state_dict(sd)
    return m.eval().cuda(), sd


def _kind(key):
    return "depth%d" % key[2] if key[0] == "depth" else key[0]


def _compare_outputs(outputs, gold, prefix, stride, worst, depth_tol=DEPTH_TOL):
    for key, val in outputs.items():
        g = gold["%s/%s" % (prefix, "_".join(str(k) for k in key))]
        got = subsample(key, val, stride).cpu().numpy()
        assert got.shape == g.shape, (key, got.shape, g.shape)
        d = float(np.abs(got - g).max())
        worst[_kind(key)] = max(worst.get(_kind(key), 0.0), d)
        assert d < (depth_tol if key[0] == "depth" else PROB_TOL), (prefix, key, d)


def _compare_state(state, gold, prefix, worst, with_key=True):
    s = FULL_STATE_STRIDE
    sv = state["values"][0][..., ::s, ::s].cpu().numpy()
    d = float(np.abs(sv - gold[prefix + "/state_value"]).max())
    worst["state_value"] = max(worst.get("state_value", 0.0), d)
    assert d < STATE_TOL, (prefix, "state_value", d)
    if with_key:
        gk = gold[prefix + "/state_key"]
        sk = state["keys"][0][..., ::s, ::s].cpu().numpy()
        d = float(np.abs(sk - gk).max() / max(1.0, np.abs(gk).max()))
        worst["state_key_rel"] = max(worst.get("state_key_rel", 0.0), d)
        assert d < STATE_TOL, (prefix, "state_key", d)


def _joint(model, height, width, gold, stride, cuda_poses=False, depth_tol=DEPTH_TOL):
    worst = {}
    state, pstate = None, None
    for w, start in enumerate((0, 3)):
        imgs, poses, K, sample = synth.synth_inputs(5, height, width, seed=0, start=start)
        if cuda_poses:
            poses, K = poses.cuda(), K.cuda()
        outputs, state, pstate = model(imgs.cuda(), poses, K, sample, state, pstate, mode="val")
        _compare_outputs(outputs, gold, "w%d" % w, stride, worst, depth_tol)
        _compare_state(state, gold, "w%d" % w, worst)
        assert np.abs(pstate[0].cpu().numpy() - gold["w%d/state_pose" % w]).max() == 0.0         # quirk Q4
    model.check()
    return worst


def test_cfg2_joint_windows_match_reference_golden():
    """The benchmark configuration itself, default arithmetic (3xf16r2d + planar feeders), host camera parameters."""
    gold = np.load(os.path.join(GOLDEN, "joint_r50_d64_480x640_g3.npz"))
    assert float(gold["meta"][6]) == synth.HEAD_GAIN
    model, _ = _model(50, 64)
    worst = _joint(model, 480, 640, gold, int(gold["meta"][4]))
    print("cfg2 480x640 D=64 R50 head_gain=3 (default arithmetic): max |diff| vs reference golden: %s" % {k: "%.2e" % v for k, v in sorted(worst.items())})
    # VERDICT r1 asked for <= 3e-4 at this setting.  Measured: 6.1e-4 in round 1; 3.6e-4 with the truncation-bias compensation on
    # the single-accumulator ring schedule; 2.0e-4 (init depth) / 1.4e-4 (fused depth) with the small products in their own
    # accumulator (the default) -- the same as the exact-fp32 kernels (2.0e-4 / 1.1e-4)
    assert max(v for k, v in worst.items() if k.startswith("depth")) < 3e-4


def test_cfg2_exact_fp32_kernels_match_reference_golden():
    """Same fixture through the exact-fp32 CUDA-core 3-D kernels and the cuDNN fp32 feeders: the floor the split arithmetic is
    measured against (two fp32 implementations with different summation orders)."""
    gold = np.load(os.path.join(GOLDEN, "joint_r50_d64_480x640_g3.npz"))
    model, _ = _model(50, 64, precision="fp32", feature_precision="fp32")
    worst = _joint(model, 480, 640, gold, int(gold["meta"][4]))
    print("cfg2 480x640 exact fp32 kernels: max |diff| vs reference golden: %s" % {k: "%.2e" % v for k, v in sorted(worst.items())})


@pytest.mark.parametrize("precision,feature_precision,depth_gate", [
    ("fp32", "fp32", 1e-3),        # floor: exact fp32 kernels + cuDNN fp32 feeders against the CPU reference
    ("3xf16", "3xf16", 1e-3),      # output-stationary tensor-core schedule (54 truncating accumulates per accumulator)
    ("3xf16r2", "3xf16", 1.5e-3),  # single-accumulator plane-ring schedule (162): see the docstring
    ("3xf16r2d", "3xf16", 1e-3),   # plane-ring schedule with the small products in a second accumulator (54)
])
def test_cfg2_head_gain_10_sweep(precision, feature_precision, depth_gate):
    """logit heads scaled by 10 (logit sigma ~ 3 over D instead of ~ 1 at the synthetic default 3).
    Depth error scales with the logit gain, so this is the stress setting of the 1e-3 gate: the exact-fp32 kernels alone use
    most of it (two fp32 implementations differ by 7.6e-4 here).  The default arithmetic (3xf16r2d: plane-ring schedule with the
    small products of the split in a second accumulator) and the output-stationary one pass it at 8.2e-4 / 8.1e-4; the
    single-accumulator ring schedule (3xf16r2, 162 truncating adds per accumulator, bias-compensated but with a three times
    larger random residual) sits at 1.25e-3 -- reported and gated at 1.5e-3, which is why it is not the default."""
    gold = np.load(os.path.join(GOLDEN, "joint_r50_d64_480x640_g10.npz"))
    assert float(gold["meta"][6]) == 10.0
    model, _ = _model(50, 64, head_gain=10.0, precision=precision, feature_precision=feature_precision)
    worst = _joint(model, 480, 640, gold, int(gold["meta"][4]), depth_tol=depth_gate)
    print("cfg2 480x640 head_gain=10, conv3d %s / feeders %s: max |diff| vs reference golden: %s"
          % (precision, feature_precision, {k: "%.2e" % v for k, v in sorted(worst.items())}))


def test_cfg3_estm_steps_match_reference_golden():
    """eval_hybrid_seq.py:169-193 at 480 x 640: 3-frame windows, 2-deep memory, stale pose (Q4), first 4 steps."""
    gold = np.load(os.path.join(GOLDEN, "estm_r50_d64_480x640.npz"))
    stride = int(gold["meta"][4])
    model, _ = _model(50, 64)
    memory, worst = [], {}
    for step in range(4):
        imgs, poses, K, sample = synth.synth_inputs(3, 480, 640, seed=0, start=step)
        pre_costs, pre_poses = sharding._flatten_memory(memory)
        outputs, costs, cposes = model(imgs.cuda(), poses, K, sample, pre_costs, pre_poses, mode="val")
        memory.append((costs, cposes))
        if len(memory) > 2:
            memory.pop(0)
        _compare_outputs(outputs, gold, "s%d" % step, stride, worst)
        _compare_state(costs, gold, "s%d" % step, worst, with_key=False)
        assert np.abs(cposes[0].cpu().numpy() - gold["s%d/state_pose" % step]).max() == 0.0
    model.check()
    print("cfg3 ESTM 480x640, 4 steps: max |diff| vs reference golden: %s" % {k: "%.2e" % v for k, v in sorted(worst.items())})


def test_cfg5_joint_windows_match_reference_golden():
    gold = np.load(os.path.join(GOLDEN, "joint_r50_d128_640x960.npz"))
    model, _ = _model(50, 128)
    worst = _joint(model, 640, 960, gold, int(gold["meta"][4]))
    print("cfg5 640x960 D=128 R50: max |diff| vs reference golden: %s" % {k: "%.2e" % v for k, v in sorted(worst.items())})


def test_cfg2_default_tf32_flags_are_not_load_bearing():
    """The drivers never touch ``cudnn.allow_tf32`` (PyTorch default: True).  The library forces strict fp32 in its cuDNN-side
    layers itself, so the fixture must still be met with the flags left at their defaults -- and they must be restored."""
    gold = np.load(os.path.join(GOLDEN, "joint_r50_d64_480x640_g3.npz"))
    before = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, True
    try:
        for feat in ("3xf16", "fp32"):
            model, _ = _model(50, 64, feature_precision=feat)
            worst = _joint(model, 480, 640, gold, int(gold["meta"][4]))
            assert torch.backends.cudnn.allow_tf32 is True and torch.backends.cuda.matmul.allow_tf32 is True
            print("cfg2 with cudnn.allow_tf32=True (caller's default), feature_precision=%s: %s" % (feat, {k: "%.2e" % v for k, v in sorted(worst.items())}))
            del model
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = before


# ------------------------------------------------------------------------------------------- CUDA camera parameters
def _gpu_geometry(poses_dev, K_dev, T, memory_poses_dev, mode):
    """The warps' matrices exactly as the model derives them from CUDA camera parameters (same functions, same device, same
    batch composition => same bits), moved to the host for the oracle's ``geometry`` hook."""
    K4 = K_dev.clone()
    K4[:, :2, :] *= 0.25
    pairs = [(t + 1, s) for t in range(T) for s in (t, t + 2)]
    homo = ops.homography_table if mode == "auto" else ops.homography_table_torch
    warp = ops.volume_warp_tables if mode == "auto" else ops.volume_warp_tables_torch
    geo = {"homo": homo(poses_dev[0].contiguous(), K4[0].contiguous(), pairs).cpu()}
    if memory_poses_dev:
        all_poses = [poses_dev[0, t + 1] for t in range(T)] + [p[0].to(torch.float32) for p in memory_poses_dev]
        geo["warp"] = [t.cpu() for t in warp(all_poses, T, K4[0].contiguous())]
    return geo


@pytest.mark.parametrize("mode", ["auto", "torch"])
@pytest.mark.parametrize("resnet,ndepths,height,width", [(18, 32, 128, 160), (50, 64, 256, 320)])
def test_cuda_poses_match_oracle_fed_the_same_matrices(resnet, ndepths, height, width, mode):
    """What eval_hybrid*.py do (``tocuda(sample)``): poses and intrinsics are CUDA tensors, so the matrices behind the warps
    are derived on the GPU -- by the library's fp64 geometry kernels (``geometry="auto"``, the default: 2 launches per window)
    or by the reference's own torch op sequence with the GPU's LU (``"torch"``).  The CPU oracle is handed exactly those
    matrices (``geometry`` hook) and computes everything downstream itself: the normal gates apply, with NO percentile
    exclusion."""
    torch.backends.cudnn.allow_tf32 = False
    model, sd = synth_model_and_state(resnet, ndepths)
    model.geometry = mode
    model.cuda()
    cfg = cfg_of(resnet, ndepths)
    T = 3
    state = pstate = ostate = opstate = None
    worst = {}
    for start in (0, 3):
        imgs, poses, K, sample = synth.synth_inputs(5, height, width, seed=0, start=start)
        poses_dev, K_dev = poses.cuda(), K.cuda()
        geo = _gpu_geometry(poses_dev, K_dev, T, pstate, mode)
        outputs, state, pstate = model(imgs.cuda(), poses_dev, K_dev, sample, state, pstate, mode="val")
        assert pstate[0].is_cuda
        with torch.no_grad():
            want, ostate, opstate = orc.forward(sd, cfg, imgs, poses, K, ostate, opstate, geometry=geo)
        for key, val in outputs.items():
            d = float((val.cpu() - want[key]).abs().max())
            worst[_kind(key)] = max(worst.get(_kind(key), 0.0), d)
            assert d < (DEPTH_TOL if key[0] == "depth" else PROB_TOL), (start, key, d)
        d = float((state["values"][0].cpu() - ostate["values"][0]).abs().max())
        worst["state_value"] = max(worst.get("state_value", 0.0), d)
        assert d < STATE_TOL, (start, d)
        assert torch.equal(pstate[0].cpu(), opstate[0])
    print("CUDA poses (geometry=%s), R%d D=%d %dx%d, oracle fed the GPU's matrices: %s" % (mode, resnet, ndepths, height, width, {k: "%.2e" % v for k, v in sorted(worst.items())}))


def test_cuda_poses_vs_reference_algorithm_on_the_same_gpu_cfg2():
    """The reference algorithm as plain PyTorch ops ON THE SAME GPU in strict fp32 (what the drivers would compute on this
    device: the GPU's LU, cuDNN fp32 convolutions, ATen grid_sample, cuBLAS for the coordinate products) against this library
    with CUDA camera parameters, at the benchmark size, both Joint windows, EVERY pixel counted, for both ways the library
    can derive the matrices on the GPU (``geometry="torch"``: the reference's op sequence with the same LU;
    ``"auto"``: the fp64 kernels, 2 launches per window).

    What differs is not arithmetic accuracy but quirk Q10: a sampling coordinate within an ulp of the +-1 cut is sampled on
    one path and zero-filled on the other.  The kernels reproduce the reference's CPU coordinate arithmetic bit for bit (the
    matrices-injected test above has no exclusions); cuBLAS orders the 3-term coordinate products differently, and another LU
    moves the matrices' last bit.  The reference shows the same thing between its OWN CPU and GPU runs, which is measured
    beside it.  Gates: "torch" (same matrices as the GPU reference): at most 1e-4 of the depth pixels at or above 1e-3;
    "auto": not more than 3x (+100) what the reference differs from itself across devices; medians are exact-fp32 class."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    models = {}
    for mode in ("torch", "auto"):
        models[mode], sd = _model(50, 64, geometry=mode)
    sd_dev = {k: v.cuda() for k, v in sd.items()}
    cfg = cfg_of(50, 64)
    states = {mode: (None, None) for mode in models}
    ostate = opstate = cstate = cpstate = None
    worst, bad, med = {m: {} for m in models}, {m: 0 for m in models}, {m: 0.0 for m in models}
    bad_ref, total = 0, 0
    torch.set_num_threads(os.cpu_count() or 1)
    for start in (0, 3):
        host = synth.synth_inputs(5, 480, 640, seed=0, start=start)
        imgs, poses, K = [t.cuda() for t in host[:3]]
        torch.set_default_device("cuda")            # the oracle's factory calls (pixel grids, plane depths) follow the inputs
        try:
            with torch.no_grad():
                want, ostate, opstate = orc.forward(sd_dev, cfg, imgs, poses, K, ostate, opstate)
        finally:
            torch.set_default_device("cpu")
        with torch.no_grad():
            cpu, cstate, cpstate = orc.forward(sd, cfg, host[0], host[1], host[2], cstate, cpstate)
        for key in want:
            if key[0] == "depth":
                bad_ref += int(((want[key].cpu() - cpu[key]).abs() >= DEPTH_TOL).sum())
                total += want[key].numel()
        for mode, model in models.items():
            outputs, st, ps = model(imgs, poses, K, None, states[mode][0], states[mode][1], mode="val")
            states[mode] = (st, ps)
            for key, val in outputs.items():
                d = (val - want[key]).abs()
                worst[mode][_kind(key)] = max(worst[mode].get(_kind(key), 0.0), float(d.max()))
                if key[0] == "depth":
                    bad[mode] += int((d >= DEPTH_TOL).sum())
                    med[mode] = max(med[mode], float(d.median()))
    for mode in models:
        print("CUDA poses (geometry=%s) vs the reference algorithm on the same GPU (strict fp32), cfg2 both windows: %s; depth pixels "
              "at/above 1e-3: %d of %d, worst median %.1e (the reference algorithm's own GPU run vs its CPU run: %d pixels)"
              % (mode, {k: "%.2e" % v for k, v in sorted(worst[mode].items())}, bad[mode], total, med[mode], bad_ref))
    assert bad["torch"] <= 1e-4 * total and med["torch"] < 1e-4, (bad, total)
    assert bad["auto"] <= 3 * bad_ref + 100 and med["auto"] < 1e-4, (bad, bad_ref, total)


# ------------------------------------------------------------------------------------------- feature cache
def test_frame_id_feature_cache_estm_cfg3():
    """ESTM at 480 x 640 with and without ``frame_ids`` (consecutive 3-frame windows share 2 frames: their matching features
    are computed once, where eval_hybrid_seq.py:169-190 recomputes them).  Same depth maps bit for bit (the kernels are batch
    invariant); the matching-feature net sees one new frame per steady-state step instead of three.  Steady-state time per step
    is printed for both (the clean-process measurement is bench.py's extras.cfg3_estm.with_frame_ids: inside a long pytest
    process the 6 ms steps are host bound and the GPU time saved does not show)."""
    model, _ = _model(50, 64)
    frames_seen = []
    hook = model.matchingFeature.register_forward_hook(lambda mod, args, out: frames_seen.append(int(args[0].shape[0])))
    torch.backends.cudnn.benchmark = False
    n_frames = 14
    windows = [synth.synth_inputs(3, 480, 640, seed=0, start=s) for s in range(n_frames - 2)]
    windows = [(w[0].cuda(), w[1], w[2]) for w in windows]

    def run(with_ids):
        memory, maps = [], []
        model._feat_cache.clear()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for s, (imgs, poses, K) in enumerate(windows):
            if s == 2:
                e0.record()                                  # steady state: two memory volumes, every window shares two frames
            pre = sharding._flatten_memory(memory)
            out, costs, cposes = model(imgs, poses, K, None, pre[0], pre[1], mode="val",
                                       frame_ids=[s, s + 1, s + 2] if with_ids else None)
            memory.append((costs, cposes))
            if len(memory) > 2:
                memory.pop(0)
            maps.append(torch.cat([out[("depth", 0, 0)], out[("depth", 0, 2)]]))
        e1.record()
        torch.cuda.synchronize()
        return torch.cat(maps), e0.elapsed_time(e1) / (len(windows) - 2)

    run(False)
    n0 = len(frames_seen)
    run(True)
    per_step_with_ids = frames_seen[n0:]
    assert frames_seen[:n0] == [3] * (n_frames - 2)                     # the reference's behaviour: every window recomputes its 3 frames
    assert per_step_with_ids == [3] + [1] * (n_frames - 3)             # with ids: 3 new frames in the first window, then 1 per step
    hook.remove()
    times = {False: [], True: []}
    for _ in range(3):
        for with_ids in (False, True):
            maps, ms = run(with_ids)
            times[with_ids].append(ms)
            if with_ids:
                cached = maps
            else:
                plain = maps
    ms_plain, ms_cached = min(times[False]), min(times[True])
    diff = float((plain - cached).abs().max())
    print("frame-id feature cache, ESTM steady state at 480x640: %.2f ms/step without ids, %.2f ms/step with ids (%.1f %% less); "
          "max |depth diff| = %.3e" % (ms_plain, ms_cached, 100.0 * (1 - ms_cached / ms_plain), diff))
    assert diff == 0.0, diff


# ------------------------------------------------------------------------------------------- driver I/O
def test_depth_map_writer_cuda_branch_is_byte_identical_and_does_not_block(tmp_path):
    """eval_hybrid.py:260-263: ``np.save(path, np.float16(outputs[key].squeeze(1).cpu().numpy()))`` -- same bytes from the
    asynchronous writer fed CUDA tensors, and ``save`` returns while the producing stream is still busy."""
    g = torch.Generator().manual_seed(0)
    maps = [(torch.rand(1, 1, 480, 640, generator=g) * 10).cuda() for _ in range(6)]
    maps[0][0, 0, 0, :4] = torch.tensor([0.1, 65504.0, 1e-8, 2049.0]).cuda()
    a = torch.randn(8192, 8192, device="cuda")
    torch.cuda.synchronize()
    busy = torch.cuda.Event()
    with DepthMapWriter() as w:
        for _ in range(20):
            a = torch.mm(a, a) * 1e-4                       # ~0.1 s of queued work ahead of the maps on the caller's stream
        outs = [m * 1.0 for m in maps]                       # the "forward" that produces the maps, behind that work
        busy.record()
        for i, o in enumerate(outs):
            w.save(o, os.path.join(tmp_path, "m%d.npy" % i))
        still_busy = not busy.query()                        # save() returned although the producer has not finished
    assert still_busy, "DepthMapWriter.save synchronised the caller's stream"
    for i, m in enumerate(maps):
        ref = os.path.join(tmp_path, "ref%d.npy" % i)
        np.save(ref, np.float16(m.squeeze(1).cpu().numpy()))
        with open(ref, "rb") as f, open(os.path.join(tmp_path, "m%d.npy" % i), "rb") as h:
            assert f.read() == h.read(), i
