"""Host-side operators over the C ABI (include/estdepth_b200.h).

Two levels:

* thin wrappers (``premix``, ``warp_cost``, ``conv3d``, ``est_attend`` ...) that take torch CUDA tensors in the
  library's native layouts (vol4 = [C/4,D,H,W,4], map4 = [C/4,H,W,4]) and enqueue one kernel on the current
  stream -- PyTorch is only the owner of the device memory here;
* reference-named operators with the reference's argument meaning and NCDHW tensors, for drop-in use and so the
  parity tests read like tests of the reference: ``homo_warping`` (utils/homo_utils.py:458), ``warp_volume``
  (utils/homo_utils.py:240), ``depthlayer`` (hybrid_models/hybrid_depth_decoder.py:33).

Nothing here falls back to PyTorch arithmetic: without the CUDA library every call raises.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvDesc, check

ACT = {"none": 0, None: 0, "relu": 1, "tanh": 2, "add_relu": 3, "sigmoid": 4}
PRECISION = {"fp32": 0, "3xtf32": 1, "3xf16": 2, "3xf16r": 3, "3xf16r2": 4, "3xf16r2d": 5}
RING_PRECISIONS = ("3xf16r", "3xf16r2", "3xf16r2d")
MAX_SOURCES = 8
# default arithmetic of conv3d: "fp32" = exact CUDA-core kernel, "3xtf32" / "3xf16" = error-compensated splits on tcgen05,
# "3xf16r" = the 3xf16 arithmetic on the plane-ring schedule (conv3d_ring.cu) for the layers it is specialised for, the
# output-stationary 3xf16 kernel for the rest
DEFAULT_PRECISION = "fp32"


class KernelProfile(object):
    """Optional per-kernel-family CUDA-event timing (bench.py's roofline pass).  Events are recorded on the stream the
    kernels are launched on; algorithmic flops / bytes per launch follow SURVEY.md section 8(d).

    Brackets (``bracket_begin`` / ``bracket_end``) time a whole stage that mixes library kernels and torch / cuDNN ops: the
    families recorded inside a bracket are remembered as its children, so that the summary can report the stage's
    EXCLUSIVE time (what is not one of the library's own families) and nothing is counted twice in the step's shares."""

    def __init__(self):
        self.records = {}          # family -> list of (start_event, end_event, flops, bytes, enclosing bracket or None)
        self.stack = []

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        return e

    def end(self, family, start, flops=0.0, nbytes=0.0):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        self.records.setdefault(family, []).append((start, e, float(flops), float(nbytes), self.stack[-1] if self.stack else None))

    def bracket_begin(self, name):
        self.stack.append(name)
        return self.begin()

    def bracket_end(self, name, start):
        assert self.stack and self.stack[-1] == name
        self.stack.pop()
        self.end(name, start)

    def summary(self, steps):
        """-> family -> (ms per step, calls per step, flops per step, bytes per step).  A bracket's entry is its EXCLUSIVE
        time; its inclusive time is reported under ``<name>(inclusive)`` with zero calls (bench.py leaves those out of the
        shares)."""
        torch.cuda.synchronize()
        out, inner = {}, {}
        for fam, recs in self.records.items():
            ms = sum(r[0].elapsed_time(r[1]) for r in recs)
            out[fam] = (ms / steps, len(recs) / float(steps), sum(r[2] for r in recs) / steps, sum(r[3] for r in recs) / steps)
            for r in recs:
                if r[4] is not None:
                    inner[r[4]] = inner.get(r[4], 0.0) + r[0].elapsed_time(r[1])
        for name, ms_inner in inner.items():
            if name in out:
                incl = out[name]
                out[name + "(inclusive)"] = (incl[0], 0.0, 0.0, 0.0)
                out[name] = (max(0.0, incl[0] - ms_inner / steps), incl[1], 0.0, 0.0)
        return out


PROFILE = None      # set to a KernelProfile() to time every launch (bench.py)


def _pb():
    return PROFILE.begin() if PROFILE is not None else None


def _pe(start, family, flops=0.0, nbytes=0.0):
    if start is not None:
        PROFILE.end(family, start, flops, nbytes)


def _bracket_begin(name):
    return PROFILE.bracket_begin(name) if PROFILE is not None else None


def _bracket_end(name, start):
    if start is not None:
        PROFILE.bracket_end(name, start)


def _stream():
    """cudaStream_t of torch's current stream on the current device, as an int for the c_void_p parameter.  (Through
    torch.cuda.current_stream() this cost 1.5 us x 2100 launches = 3 ms of host time per step, profiles/host_profile.py.)"""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _ptr(t, dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("estdepth_b200 ops need CUDA tensors (there is no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError("estdepth_b200 ops need contiguous %s tensors, got %s contiguous=%s" %
                           (dtype, t.dtype, t.is_contiguous()))
    return ctypes.c_void_p(t.data_ptr())


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------ thin wrappers

def homography_setup(ref_pose, src_pose, cam_intr, out=None):
    """[4,4], [4,4], [3,3] device tensors -> [12] = rot(9) | trans(3)."""
    out = torch.empty(12, device=ref_pose.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_homography_setup(_ptr(ref_pose), _ptr(src_pose), _ptr(cam_intr), _ptr(out), _stream()),
          "estd_homography_setup")
    _pe(t, "geometry_setup")
    return out


def homography_from_proj(src_proj, ref_proj, out=None):
    out = torch.empty(12, device=src_proj.device, dtype=torch.float32) if out is None else out
    check(_lib.get().estd_homography_from_proj(_ptr(src_proj), _ptr(ref_proj), _ptr(out), _stream()),
          "estd_homography_from_proj")
    return out


def volume_warp_setup(pose_i, pose_j, cam_intr, out=None):
    """-> [30] = Kinv(9) | Minv 3x4 (12) | K(9), Minv = (P_j P_i^-1)^-1."""
    out = torch.empty(30, device=pose_i.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_volume_warp_setup(_ptr(pose_i), _ptr(pose_j), _ptr(cam_intr), _ptr(out), _stream()),
          "estd_volume_warp_setup")
    _pe(t, "geometry_setup")
    return out


def homography_table(poses, K4, pairs, out=None):
    """[rot | trans] of every (reference view, source view) pair in ONE launch (fp64 algebra, rounded once):
    poses [V,4,4], K4 [3,3] device tensors, pairs = [(ref, src), ...] -> [len(pairs), 12]."""
    n = len(pairs)
    out = torch.empty(n, 12, device=poses.device, dtype=torch.float32) if out is None else out
    flat = (ctypes.c_int * (2 * n))(*[int(i) for pr in pairs for i in pr])
    t = _pb()
    check(_lib.get().estd_homography_table(_ptr(poses), poses.shape[0], _ptr(K4), flat, n, _ptr(out), _stream()), "estd_homography_table")
    _pe(t, "geometry_setup")
    return out


def volume_warp_tables(all_poses, n_targets, K4):
    """[Kinv | Minv (3x4) | K] for every (target i < n_targets, source j != i) pair of ``all_poses`` (list of [4,4] device
    tensors) in ONE launch -> list (per target) of [len(all_poses) - 1, 30], sources in list order."""
    n = len(all_poses)
    pairs = [(i, j) for i in range(n_targets) for j in range(n) if j != i]
    m = len(pairs)
    out = torch.empty(m, 30, device=K4.device, dtype=torch.float32)
    ptrs = (ctypes.c_void_p * n)(*[_ptr(p).value for p in all_poses])
    flat = (ctypes.c_int * (2 * m))(*[int(i) for pr in pairs for i in pr])
    t = _pb()
    check(_lib.get().estd_volume_warp_table(ptrs, n, _ptr(K4), flat, m, _ptr(out), _stream()), "estd_volume_warp_table")
    _pe(t, "geometry_setup")
    return [out[i * (n - 1):(i + 1) * (n - 1)] for i in range(n_targets)]


_INV_MIN_BATCH = 8


def _inv(m):
    """torch.inverse without blocking the host.  ``torch.inverse`` checks its result (one synchronisation per call), and
    even ``linalg.inv_ex`` synchronises when it is given a SINGLE 4x4 matrix (looped cuSOLVER path; measured in
    profiles/inv_sync.py) -- the batched path does not, and gives the same bits per matrix (profiles/inv_bits.py).  So the
    batch is padded with identities to at least 8 matrices."""
    n = m.shape[0]
    if n < _INV_MIN_BATCH:
        pad = torch.eye(m.shape[-1], device=m.device, dtype=m.dtype).expand(_INV_MIN_BATCH - n, -1, -1)
        return torch.linalg.inv_ex(torch.cat([m, pad], 0))[0][:n]
    return torch.linalg.inv_ex(m)[0]


def homography_table_torch(poses, K4, pairs):
    """The [rot | trans] of every (reference view, source view) pair computed with the REFERENCE'S OWN fp32 torch ops, in its
    order (model_hybrid.py:74-88, homo_utils.py:469-471): ext = inverse(pose); proj[:3,:4] = K ext[:3,:4];
    M = src_proj inverse(ref_proj).  poses [V,4,4], K4 [3,3], pairs = [(ref, src), ...] -> [len(pairs), 12].

    Why not the fp64 kernel (estd_homography_setup): a sampling coordinate within an ulp of the [-1, 1] range flips between
    'sampled' and 'zero-filled' (quirk Q10) when the matrices differ in the last bit; with the reference's matrices the
    kernels' coordinate arithmetic reproduces the reference's bit for bit (tests/run_fullsize_parity.py)."""
    V = poses.shape[0]
    ext = _inv(poses)                                                    # [V,4,4]
    proj = []
    for v in range(V):
        p = ext[v:v + 1].clone()
        p[:, :3, :4] = K4.unsqueeze(0) @ ext[v:v + 1, :3, :4]            # batch-1 bmm, as the reference issues it
        proj.append(p)
    refs = sorted(set(r for r, _ in pairs))
    ref_inv = _inv(torch.cat([proj[r] for r in refs], 0))
    rows = []
    for r, s_ in pairs:
        m = torch.matmul(proj[s_], ref_inv[refs.index(r):refs.index(r) + 1])[0]
        rows.append(torch.cat([m[:3, :3].reshape(9), m[:3, 3]]))
    return torch.stack(rows).contiguous()


def volume_warp_tables_torch(all_poses, n_targets, K4):
    """[Kinv | Minv (3x4) | K] for every (target i < n_targets, source j != i) pair of ``all_poses`` with the reference's fp32
    torch ops (hybrid_depth_decoder.py:235: rel = P_j inverse(P_i); homo_utils.py:51,258: inverse(K), inverse(rel)).
    all_poses: list of [4,4]; K4 [3,3] -> list (per target) of [len(all_poses) - 1, 30], sources in list order."""
    n = len(all_poses)
    inv_t = _inv(torch.stack(all_poses[:n_targets]))                     # one batched LU: same bits as one at a time
    rel = torch.cat([all_poses[j].unsqueeze(0) @ inv_t[i:i + 1]          # batch-1 matmuls, as the reference issues them
                     for i in range(n_targets) for j in range(n) if j != i], 0)
    m = n_targets * (n - 1)
    minv = _inv(rel)[:, :3, :4].reshape(m, 12)
    kinv = _inv(K4.unsqueeze(0)).reshape(1, 9).expand(m, 9)
    table = torch.cat([kinv, minv, K4.reshape(1, 9).expand(m, 9)], 1).contiguous()
    return [table[i * (n - 1):(i + 1) * (n - 1)] for i in range(n_targets)]


def premix(fea_chw, weight, bias=None, out=None):
    """fea [Cin,H,W], weight [Cout,Cin], bias [Cout]|None -> map4 [Cout/4,H,W,4]."""
    cin, H, W = fea_chw.shape
    cout = weight.shape[0]
    out = torch.empty(cout // 4, H, W, 4, device=fea_chw.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_premix(_ptr(fea_chw), _ptr(weight), _ptr(bias), _ptr(out), cin, cout, H, W, _stream()),
          "estd_premix")
    _pe(t, "premix", 2.0 * cin * cout * H * W, 4.0 * (cin + cout) * H * W)
    return out


def premix_batch(fea_nchw, weight, bias=None, out=None):
    """fea [N,Cin,H,W], weight [Cout,Cin], bias [Cout]|None -> [N,Cout/4,H,W,4] (N map4 tensors, one launch)."""
    n, cin, H, W = fea_nchw.shape
    cout = weight.shape[0]
    out = torch.empty(n, cout // 4, H, W, 4, device=fea_nchw.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_premix_batch(_ptr(fea_nchw), _ptr(weight), _ptr(bias), _ptr(out), n, cin, cout, H, W, _stream()),
          "estd_premix_batch")
    _pe(t, "premix", 2.0 * n * cin * cout * H * W, 4.0 * n * (cin + cout) * H * W)
    return out


def warp_cost(ref_mix, src_mix, homo12, depth_values, out=None, align_corners=False):
    """map4 [C/4,H,W,4] x2, [12], [D] -> vol4 [C/4,D,H,W,4]."""
    chunks, H, W, _ = ref_mix.shape
    D = depth_values.numel()
    out = torch.empty(chunks, D, H, W, 4, device=ref_mix.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_warp_cost(_ptr(ref_mix), _ptr(src_mix), _ptr(homo12), _ptr(depth_values), _ptr(out),
                                    chunks * 4, D, H, W, int(bool(align_corners)), _stream()), "estd_warp_cost")
    _pe(t, "warp_cost", 0.0, 4.0 * chunks * 4 * H * W * (D + 2))       # 4*C*P*(D+2)
    return out


class PackedConv(object):
    """Folded, packed parameters of one 3x3x3 layer (see packing.pack_conv3d)."""
    __slots__ = ("weight", "scale", "shift", "cin_chunks", "cout_pad", "out_chunks", "act_split", "act_lo", "act_hi",
                 "cin", "cout", "weight_tc", "cout_pad_tc", "weight_f16", "scale_f16", "weight_ring", "weight_ring2", "scale_ring", "_desc",
                 "precision", "cout_pad_ring2", "weight_ring2d")

    def __init__(self, weight, scale, shift, cin_chunks, cout_pad, out_chunks, act_split, act_lo, act_hi,
                 cin=None, cout=None, weight_tc=None, cout_pad_tc=None):
        self.weight, self.scale, self.shift = weight, scale, shift
        self._desc = None                                              # per-(arithmetic, mode) descriptor templates (_conv_desc)
        self.precision = None                                          # per-layer arithmetic override (None: the caller's choice)
        self.cout_pad_ring2 = None                                     # columns per ring slot of a narrow layer's CTA-pair packing
        self.weight_ring2d = None                                      # CTA-pair packing for two accumulators per slot (merged B operand)
        self.weight_tc, self.cout_pad_tc = weight_tc, cout_pad_tc      # tcgen05 packings (packing.attach_tc)
        self.weight_f16, self.scale_f16 = None, None
        self.weight_ring = None                                        # plane-ring packing (packing.pack_weight_ring)
        self.weight_ring2 = None                                       # CTA-pair plane-ring packing (packing.pack_weight_ring2)
        self.scale_ring = None                                         # uniform 2^-k multiplier of the ring packings
        self.cin = cin if cin is not None else 4 * cin_chunks          # real (un-padded) channel counts, for flop accounting
        self.cout = cout if cout is not None else min(cout_pad, 4 * out_chunks)
        self.cin_chunks, self.cout_pad, self.out_chunks = cin_chunks, cout_pad, out_chunks
        self.act_split, self.act_lo, self.act_hi = act_split, ACT[act_lo], ACT[act_hi]

    TENSORS = ("weight", "scale", "shift", "weight_tc", "weight_f16", "scale_f16", "weight_ring", "weight_ring2", "scale_ring",
               "weight_ring2d")

    def to(self, device, memo=None):
        """Moves every buffer to ``device`` in place (one copy per buffer; ``memo`` de-duplicates buffers shared between
        layers) and drops the cached descriptor templates, which hold device addresses.  Returns self."""
        device = torch.device(device)
        for name in self.TENSORS:
            t = getattr(self, name)
            if t is None or t.device == device:
                continue
            if memo is not None:
                got = memo.get(id(t))
                if got is None:
                    got = memo[id(t)] = (t, t.to(device))
                setattr(self, name, got[1])
            else:
                setattr(self, name, t.to(device))
        self._desc = None
        return self


def _precision(pc, precision):
    precision = DEFAULT_PRECISION if precision is None else precision
    if precision not in PRECISION:
        raise RuntimeError("conv3d: unknown precision %r" % (precision,))
    if precision == "3xf16r2d" and pc.weight_ring2d is None:
        precision = "3xf16r2"          # two accumulators per slot do not fit TMEM for this shape
    if precision == "3xf16r2" and pc.weight_ring2 is None:
        precision = "3xf16r"           # same arithmetic and schedule on single CTAs: no CTA-pair specialisation for this shape
    if precision == "3xf16r" and pc.weight_ring is None:
        precision = "3xf16"            # same arithmetic, output-stationary schedule: no ring specialisation for this shape
    if (precision == "3xtf32" and pc.weight_tc is None) or (precision == "3xf16" and pc.weight_f16 is None):
        raise RuntimeError("conv3d: layer was packed without tensor-core weights (packing.attach_tc)")
    return precision


_STATUS = {}


def status_flag(device):
    """Per-device int32 flag the fp16-split kernels raise when an activation leaves the fp16 range."""
    key = str(device)
    if key not in _STATUS:
        _STATUS[key] = torch.zeros(1, device=device, dtype=torch.int32)
    return _STATUS[key]


def check_status(device):
    """Raises if any 3xf16 convolution saw |x| > 65504 since the last check (one 4-byte D2H read; synchronises)."""
    flag = _STATUS.get(str(device))
    if flag is not None and int(flag.item()) != 0:
        flag.zero_()
        raise RuntimeError("estdepth_b200: an activation exceeded the fp16 range in a 3xf16 convolution; "
                           "results are invalid -- use precision='3xtf32' or 'fp32' for this model")


_STATUS_ASYNC = {}


def check_status_async(device):
    """The same check WITHOUT blocking the host: the flag is copied to pinned memory behind the work already enqueued, and
    the copy of an EARLIER call is examined once its event has completed.  A violation is therefore reported one or two
    calls late -- but a model's forward no longer drains the GPU before it starts issuing (the blocking check cost the
    whole issue latency of the first kernels of every step)."""
    key = str(device)
    flag = _STATUS.get(key)
    if flag is None:
        return
    ent = _STATUS_ASYNC.get(key)
    if ent is not None and ent[1].query():
        if int(ent[0][0]) != 0:
            flag.zero_()
            _STATUS_ASYNC.pop(key, None)
            raise RuntimeError("estdepth_b200: an activation exceeded the fp16 range in a 3xf16 convolution of an earlier "
                               "call; its results are invalid -- use precision='3xtf32' or 'fp32' for this model")
        ent = None
    if ent is None:
        host = torch.empty(1, dtype=torch.int32, device="cpu").pin_memory()
        host.copy_(flag, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(flag.device))
        _STATUS_ASYNC[key] = (host, ev)


def _act_ptr(t, dtype=torch.float32):
    """Device address of an activation tensor as a plain int (c_void_p fields take ints), with _ptr's checks."""
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        _ptr(t, dtype)                     # raises with the full message
    return t.data_ptr()


def _desc_template(pc, precision, planar, dilation, device):
    """The layer-constant half of an estd_conv3d_desc (weights, affine arrays, activation codes, arithmetic), built once per
    (layer, arithmetic, mode, device) and byte-copied per launch: filling a ctypes structure field by field costs more host
    time than everything else a launch does.  The tensors the pointers refer to stay alive through ``pc`` / status_flag."""
    d = ConvDesc()
    d.precision = PRECISION[precision]
    d.planar, d.dilation = int(planar), int(dilation)
    tc = precision != "fp32"
    d.weight = _ptr(pc.weight)
    f16 = precision in ("3xf16",) + RING_PRECISIONS
    d.weight_tc = _ptr(pc.weight_ring2d if precision == "3xf16r2d" else pc.weight_ring2 if precision == "3xf16r2" else pc.weight_ring if precision == "3xf16r"
                       else pc.weight_f16 if precision == "3xf16" else pc.weight_tc)
    ring = precision in RING_PRECISIONS
    d.scale, d.shift = _ptr(pc.scale_ring if ring else pc.scale_f16 if f16 else pc.scale), _ptr(pc.shift)
    d.cout_pad = pc.cout_pad_tc if tc else pc.cout_pad
    if precision in ("3xf16r2", "3xf16r2d") and pc.cout_pad_ring2 is not None:
        d.cout_pad = pc.cout_pad_ring2
    d.status = _ptr(status_flag(device), torch.int32) if f16 else None
    d.act_split, d.act_lo, d.act_hi = pc.act_split, pc.act_lo, pc.act_hi
    return bytes(d)


def _even(n):
    return (n + 1) // 2 * 2


def _conv_desc(pc, in0, in1, out0, out1, res0, res1, post_scale, gn_partials, precision, planar=0, dilation=1,
               in_split=(False, False), res_split=False, out_split=False, head=None, out_up2=False):
    chunks0, D, H, W, _ = in0.shape
    key = (precision, planar, dilation, in0.device.index)
    cache = pc._desc
    if cache is None:
        cache = pc._desc = {}
    tmpl = cache.get(key)
    if tmpl is None:
        tmpl = cache[key] = _desc_template(pc, precision, planar, dilation, in0.device)
    d = ConvDesc.from_buffer_copy(tmpl)
    d.in0, d.in0_chunks = _act_ptr(in0), chunks0
    if in1 is not None:
        d.in1, d.in1_chunks = _act_ptr(in1), in1.shape[0]
    # a pre-split (vol4s) tensor has an even number of chunks: a 36-channel tensor occupies 10
    want_in = _even(pc.cin_chunks) if (in_split[0] and in1 is None) else pc.cin_chunks
    if d.in0_chunks + d.in1_chunks != want_in:
        raise RuntimeError("conv3d: layer packed for %d input chunks, got %d" % (want_in, d.in0_chunks + d.in1_chunks))
    if any(in_split) or res_split or out_split or head is not None:
        if precision not in RING_PRECISIONS and not planar:
            raise RuntimeError("conv3d: pre-split tensors / the fused logit head need the plane-ring kernels, not %r" % (precision,))
        if (in_split[0] and chunks0 % 2) or (in1 is not None and in_split[1] and in1.shape[0] % 2) or (out_split and out1 is not None):
            raise RuntimeError("conv3d: pre-split tensors hold an even number of chunks; a split output has one segment")
        d.in0_split, d.in1_split, d.res_split, d.out_split = int(in_split[0]), int(in_split[1]), int(res_split), int(out_split)
    d.res0, d.res1 = _act_ptr(res0), _act_ptr(res1)
    d.post_scale = post_scale
    if out0 is not None:
        d.out0, d.out0_chunks = _act_ptr(out0), out0.shape[0]
    else:
        d.out0_chunks = pc.out_chunks
    if out1 is not None:
        d.out1, d.out1_chunks = _act_ptr(out1), out1.shape[0]
    if d.out0_chunks + d.out1_chunks != (_even(pc.out_chunks) if out_split else pc.out_chunks):
        raise RuntimeError("conv3d: layer produces %d chunks, outputs hold %d" % (pc.out_chunks, d.out0_chunks + d.out1_chunks))
    if out_up2:
        if not planar or out1 is not None or tuple(out0.shape[1:]) != (D, 2 * H, 2 * W, 4):
            raise RuntimeError("conv: an up-sampled output needs a planar layer and one [chunks, N, 2H, 2W, 4] tensor")
        d.out_up2 = 1
    if head is not None:
        if planar or pc.cout_pad_tc != 16 or gn_partials is not None:
            raise RuntimeError("conv3d: the fused logit head needs a 16-channel 3x3x3 layer")
        d.head_w, d.head_b, d.head_out = _act_ptr(head[0]), _act_ptr(head[1]), _act_ptr(head[2])
    elif out0 is None:
        raise RuntimeError("conv3d: no output")
    if gn_partials is not None:
        d.gn_partials = _act_ptr(gn_partials, torch.float64)
    d.D, d.H, d.W = D, H, W
    return d


def conv3d_num_ctas(pc, D, H, W, precision=None):
    precision = _precision(pc, pc.precision or precision)
    d = ConvDesc()
    d.precision = PRECISION[precision]
    d.in0_chunks, d.in1_chunks = pc.cin_chunks, 0
    d.cout_pad = pc.cout_pad_tc if precision != "fp32" else pc.cout_pad
    if precision in ("3xf16r2", "3xf16r2d") and pc.cout_pad_ring2 is not None:
        d.cout_pad = pc.cout_pad_ring2
    d.D, d.H, d.W = D, H, W
    n = _lib.get().estd_conv3d_num_ctas(ctypes.byref(d))
    if n < 0:
        check(n, "estd_conv3d_num_ctas")
    return n


def conv3d(pc, in0, out0, in1=None, out1=None, res0=None, res1=None, post_scale=1.0, gn_partials=None, precision=None,
           in_split=(False, False), res_split=False, out_split=False, head=None):
    """3x3x3 conv + folded affine + activation (+ residuals, x post_scale) over vol4 tensors; returns out0.

    precision: "fp32" (exact, CUDA cores) | "3xtf32" | "3xf16" (tcgen05 tensor cores, error-compensated splits) |
    "3xf16r" / "3xf16r2" (the same arithmetic on the plane-ring schedules) | None = DEFAULT_PRECISION.
    Plane-ring kernels only: ``in_split`` = (in0, in1) are pre-split (vol4s, ``to_split``), ``res_split`` = the residuals
    are, ``out_split`` = write out0 pre-split; ``head`` = (weight [16], bias [1], logits_out [D,H,W]): the 1x1x1 logit head
    fused into the epilogue of a 16-channel layer (out0 may then be None)."""
    precision = _precision(pc, pc.precision or precision)
    d = _conv_desc(pc, in0, in1, out0, out1, res0, res1, post_scale, gn_partials, precision,
                   in_split=in_split, res_split=res_split, out_split=out_split, head=head)
    t = _pb()
    check(_lib.get().estd_conv3d(ctypes.byref(d), _stream()), "estd_conv3d")
    vox = float(d.D) * d.H * d.W
    _pe(t, "conv3d_" + precision, 54.0 * pc.cin * pc.cout * vox, 4.0 * vox * (pc.cin + pc.cout))
    return out0


def conv_planar(pc, in0, out0, res0=None, dilation=1, in1=None, taps=9, post_scale=1.0,
                in_split=(False, False), res_split=False, out_split=False, out_up2=False):
    """2-D 3x3 (taps=9; stride 1, padding = dilation) or 1x1 (taps=1) convolution over a stack of N maps held as vol4
    [C/4,N,H,W,4], with folded affine + activation (+ residual), on the tensor cores (fp16 two-term split).  Returns out0.
    ``in_split`` / ``res_split`` / ``out_split``: pre-split (vol4s) tensors, as for ``conv3d``; ``out_up2``: out0 is
    [C/4,N,2H,2W,4] and receives the result nearest-neighbour x2 up-sampled."""
    d = _conv_desc(pc, in0, in1, out0, None, res0, None, post_scale, None, "3xf16", planar=1 if taps == 9 else 2, dilation=dilation,
                   in_split=in_split, res_split=res_split, out_split=out_split, out_up2=out_up2)
    t = _pb()
    check(_lib.get().estd_conv3d(ctypes.byref(d), _stream()), "estd_conv3d(planar)")
    vox = float(d.D) * d.H * d.W
    _pe(t, "conv2d_3xf16", 2.0 * taps * pc.cin * pc.cout * vox, 4.0 * vox * (pc.cin + pc.cout))
    return out0


def to_split(vol4):
    """vol4 -> vol4s with torch ops (tests, one-off conversions): chunks (2g, 2g+1) <- (fp16(x), fp16(x - fp16(x))) of channels
    8g..8g+7, bit-identical to what the kernels' splitter warps / split epilogues produce.  The chunk count is padded to even."""
    chunks = vol4.shape[0]
    if chunks % 2:
        vol4 = torch.cat([vol4, torch.zeros_like(vol4[:1])], 0)
        chunks += 1
    x = vol4.reshape(chunks // 2, 2, *vol4.shape[1:-1], 4)                    # [g, half, ..., 4]
    x = torch.cat([x[:, 0], x[:, 1]], dim=-1)                                 # [g, ..., 8 channels]
    hi = x.to(torch.float16)
    lo = (x - hi.to(torch.float32)).to(torch.float16)
    out = torch.stack([hi, lo], dim=1)                                        # [g, 2, ..., 8] fp16
    return out.reshape(chunks, *vol4.shape[1:-1], 8).contiguous().view(torch.float32)


def from_split(vol4s):
    """vol4s -> vol4 (x_hi + x_lo)."""
    chunks = vol4s.shape[0]
    h = vol4s.contiguous().view(torch.float16).reshape(chunks // 2, 2, *vol4s.shape[1:-1], 8).to(torch.float32)
    x = h[:, 0] + h[:, 1]                                                     # [g, ..., 8]
    x = torch.stack([x[..., :4], x[..., 4:]], dim=1)                          # [g, 2, ..., 4]
    return x.reshape(chunks, *vol4s.shape[1:-1], 4).contiguous()


def nchw_to_vol4(x, out=None):
    """[N,C,H,W] -> vol4 [C/4,N,H,W,4] (``out`` may be a chunk slice of a wider vol4 buffer)."""
    N, C, H, W = x.shape
    out = torch.empty(C // 4, N, H, W, 4, device=x.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_nchw_to_vol4(_ptr(x), _ptr(out), N, C, H, W, _stream()), "estd_nchw_to_vol4")
    _pe(t, "layout", 0.0, 8.0 * N * C * H * W)
    return out


def stem_conv(img_nchw, weight, bias, out=None, out_split=False):
    """Conv2d(3, 32, 3, stride 2, pad 1) + folded affine + ReLU: NCHW images [N,3,H,W] -> vol4 [8,N,Ho,Wo,4] (optionally pre-split)."""
    N, C, H, W = img_nchw.shape
    if C != 3 or tuple(weight.shape) != (32, 3, 3, 3):
        raise RuntimeError("stem_conv: 3 -> 32 channels, 3x3 filter only")
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(8, N, Ho, Wo, 4, device=img_nchw.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_stem_conv(_ptr(img_nchw), _ptr(weight), _ptr(bias), _ptr(out), N, H, W, int(bool(out_split)),
                                    _ptr(status_flag(img_nchw.device), torch.int32) if out_split else None, _stream()), "estd_stem_conv")
    _pe(t, "stem_conv", 2.0 * 27 * 32 * N * Ho * Wo, 4.0 * N * (3 * H * W + 32 * Ho * Wo))
    return out


def stem7_conv(img_nchw, weight, bias, out=None, out_split=False):
    """torchvision ResNet conv1: Conv2d(3, 64, 7, stride 2, pad 3) + folded affine + ReLU: NCHW images [N,3,H,W] -> vol4
    [16,N,Ho,Wo,4] (optionally pre-split).  ``weight`` is TAP-MAJOR: the PyTorch weight [64,3,7,7] permuted to [3,7,7,64]."""
    N, C, H, W = img_nchw.shape
    if C != 3 or tuple(weight.shape) != (3, 7, 7, 64):
        raise RuntimeError("stem7_conv: 3 -> 64 channels, 7x7 filter, weight permuted to [3, 7, 7, 64]")
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(16, N, Ho, Wo, 4, device=img_nchw.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_stem7_conv(_ptr(img_nchw), _ptr(weight), _ptr(bias), _ptr(out), N, H, W, int(bool(out_split)),
                                     _ptr(status_flag(img_nchw.device), torch.int32) if out_split else None, _stream()), "estd_stem7_conv")
    _pe(t, "stem_conv", 2.0 * 147 * 64 * N * Ho * Wo, 4.0 * N * (3 * H * W + 64 * Ho * Wo))
    return out


def maxpool3x3s2_vol4(x4, out=None, in_split=False, out_split=False):
    """MaxPool2d(3, stride 2, pad 1) over vol4 maps [C/4,N,H,W,4] -> [C/4,N,Ho,Wo,4]; either side may be pre-split (vol4s)."""
    chunks, N, H, W, _ = x4.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(chunks, N, Ho, Wo, 4, device=x4.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_maxpool3x3s2_vol4(_ptr(x4), _ptr(out), chunks, N, H, W, int(bool(in_split)), int(bool(out_split)),
                                            _ptr(status_flag(x4.device), torch.int32) if out_split else None, _stream()), "estd_maxpool3x3s2_vol4")
    _pe(t, "layout", 0.0, 16.0 * chunks * N * (H * W + Ho * Wo))
    return out


def vol4_to_nchw(v, out=None):
    chunks, N, H, W, _ = v.shape
    out = torch.empty(N, chunks * 4, H, W, device=v.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_vol4_to_nchw(_ptr(v), _ptr(out), N, chunks * 4, H, W, _stream()), "estd_vol4_to_nchw")
    _pe(t, "layout", 0.0, 8.0 * N * chunks * 4 * H * W)
    return out


def upsample_bilinear_vol4(src, out, bias=None, relu=False):
    """F.interpolate(relu?(src + bias), size=out's, mode='bilinear', align_corners=False) of NCHW maps [N,C,h,w], written as
    vol4 [C/4,N,H,W,4] into ``out`` (which may be a chunk slice of a wider vol4 buffer)."""
    N, C, h, w = src.shape
    chunks, N2, H, W, _ = out.shape
    if chunks * 4 != C or N2 != N:
        raise RuntimeError("upsample_bilinear_vol4: %s does not hold %s" % (tuple(out.shape), tuple(src.shape)))
    t = _pb()
    check(_lib.get().estd_upsample_bilinear_vol4(_ptr(src), _ptr(bias), _ptr(out), N, C, h, w, H, W, int(bool(relu)), _stream()),
          "estd_upsample_bilinear_vol4")
    _pe(t, "layout", 0.0, 4.0 * N * C * (h * w + H * W))
    return out


def est_attend(key_t, src_keys, src_values, warp30, depth_values, depth_min, depth_interval, out=None,
               align_corners=False):
    """key_t vol4 [4,D,H,W,4]; lists of N source key/value vol4; warp30 [N,30] -> h vol4 [4,D,H,W,4]."""
    n = len(src_keys)
    if n < 1 or n > MAX_SOURCES or len(src_values) != n:
        raise RuntimeError("est_attend: need 1..%d sources, got %d" % (MAX_SOURCES, n))
    _, D, H, W, _ = key_t.shape
    out = torch.empty_like(key_t) if out is None else out
    karr = (ctypes.c_void_p * n)(*[_ptr(k).value for k in src_keys])
    varr = (ctypes.c_void_p * n)(*[_ptr(v).value for v in src_values])
    t = _pb()
    check(_lib.get().estd_est_attend(_ptr(key_t), n, karr, varr, _ptr(warp30), _ptr(depth_values), float(depth_min),
                                     float(depth_interval), _ptr(out), D, H, W, int(bool(align_corners)), _stream()),
          "estd_est_attend")
    _pe(t, "est_attend", 0.0, 4.0 * 16 * D * H * W * (2 + 2 * n))
    return out


def gn_finalize(partials, n_groups, count_per_group, eps=1e-5, out=None):
    out = torch.empty(4, device=partials.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_gn_finalize(_ptr(partials, torch.float64), partials.shape[0], n_groups, float(count_per_group),
                                      float(eps), _ptr(out), _stream()), "estd_gn_finalize")
    _pe(t, "gn_finalize")
    return out


def gru_reset(f, h, stats, gamma, beta, out=None):
    _, D, H, W, _ = h.shape
    out = torch.empty_like(h) if out is None else out
    t = _pb()
    check(_lib.get().estd_gru_reset(_ptr(f), _ptr(h), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(out), D, H, W, _stream()),
          "estd_gru_reset")
    _pe(t, "gru_reset", 0.0, 4.0 * 16 * D * H * W * 3)
    return out


def gru_blend(f, h, o, stats_f, stats_o, gamma_u, beta_u, gamma_o, beta_o, out=None):
    _, D, H, W, _ = h.shape
    out = torch.empty_like(h) if out is None else out
    t = _pb()
    check(_lib.get().estd_gru_blend(_ptr(f), _ptr(h), _ptr(o), _ptr(stats_f), _ptr(stats_o), _ptr(gamma_u), _ptr(beta_u),
                                    _ptr(gamma_o), _ptr(beta_o), _ptr(out), D, H, W, _stream()), "estd_gru_blend")
    _pe(t, "gru_blend", 0.0, 4.0 * 16 * D * H * W * 4)
    return out


def head_softargmin(depth_values, hidden=None, head_w=None, head_b=None, logits_in=None, logits_out=None,
                    depth_out=None, prob_out=None, argmax_out=None, up=4, shape=None):
    """Logit head (optional) + softmax over D + expectation; outputs are written `up` x replicated."""
    if hidden is not None:
        _, D, H, W, _ = hidden.shape
    else:
        D, H, W = logits_in.shape
    t = _pb()
    check(_lib.get().estd_head_softargmin(_ptr(hidden), _ptr(head_w), _ptr(head_b), _ptr(logits_in), _ptr(depth_values),
                                          _ptr(logits_out), _ptr(depth_out), _ptr(prob_out),
                                          _ptr(argmax_out, torch.int32), D, H, W, up, _stream()), "estd_head_softargmin")
    n_out = sum(x is not None for x in (depth_out, prob_out, argmax_out))
    _pe(t, "head_softargmin", 0.0, 4.0 * ((16 if hidden is not None else 1) * D * H * W
                                          + (D * H * W if logits_out is not None else 0) + n_out * up * up * H * W))


def vol4_to_ncdhw(vol4, out=None):
    chunks, D, H, W, _ = vol4.shape
    out = torch.empty(chunks * 4, D, H, W, device=vol4.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_vol4_to_ncdhw(_ptr(vol4), _ptr(out), chunks * 4, D, H, W, _stream()), "estd_vol4_to_ncdhw")
    _pe(t, "layout", 0.0, 8.0 * chunks * 4 * D * H * W)
    return out


def ncdhw_to_vol4(x, out=None):
    C, D, H, W = x.shape
    out = torch.empty(C // 4, D, H, W, 4, device=x.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_ncdhw_to_vol4(_ptr(x), _ptr(out), C, D, H, W, _stream()), "estd_ncdhw_to_vol4")
    _pe(t, "layout", 0.0, 8.0 * C * D * H * W)
    return out


def scalar_to_vol4(x, out=None):
    D, H, W = x.shape
    out = torch.empty(1, D, H, W, 4, device=x.device, dtype=torch.float32) if out is None else out
    t = _pb()
    check(_lib.get().estd_scalar_to_vol4(_ptr(x), _ptr(out), D, H, W, _stream()), "estd_scalar_to_vol4")
    _pe(t, "layout", 0.0, 20.0 * D * H * W)
    return out


# ------------------------------------------------------------------------------------ reference-named operators

def homo_warping(src_fea, src_proj, ref_proj, depth_values, align_corners=False):
    """Same contract as the reference's ``homo_warping`` (utils/homo_utils.py:458-504).

    src_fea [B,C,H,W], src_proj/ref_proj [B,4,4], depth_values [B,D] or [B,D,1,1] -> [B,C,D,H,W].
    (The product path never calls this un-fused form; it exists as the operator-level drop-in.)
    """
    B, C, H, W = src_fea.shape
    dv = _f32c(depth_values).reshape(B, -1)
    D = dv.shape[1]
    Cp = (C + 3) // 4 * 4
    eye = torch.eye(Cp, C, device=src_fea.device, dtype=torch.float32)
    out = torch.empty(B, Cp, D, H, W, device=src_fea.device, dtype=torch.float32)
    zeros = torch.zeros(Cp // 4, H, W, 4, device=src_fea.device, dtype=torch.float32)
    for b in range(B):
        src_mix = premix(_f32c(src_fea[b]), eye)
        h12 = homography_from_proj(_f32c(src_proj[b]), _f32c(ref_proj[b]))
        vol = warp_cost(zeros, src_mix, h12, dv[b].contiguous(), align_corners=align_corners)
        vol4_to_ncdhw(vol, out[b])
    return out[:, :C]


def warp_volume(feat_volume, depth, pose, cam_intr, pixel_coords=None, depth_min=None, depth_interval=None,
                align_corners=False):
    """Same contract as the reference's ``warp_volume`` (utils/homo_utils.py:240-279), zeros padding.

    feat_volume [N,C,D,H,W] (C <= 16), depth [N,1,D,H*W] (the plane depths, constant over H*W as in the decoder,
    hybrid_depth_decoder.py:237), pose [N,4,4] = P_j P_i^-1, cam_intr [N,3,3].  ``pixel_coords`` is accepted and
    ignored (the kernel generates the pixel grid).  Implemented as the fused EST gather with a single source,
    whose softmax weight is exactly 1.
    """
    N, C, D, H, W = feat_volume.shape
    if C > 16:
        raise RuntimeError("warp_volume: at most 16 channels (the EST key/value width)")
    dev = feat_volume.device
    out = torch.empty(N, 16, D, H, W, device=dev, dtype=torch.float32)
    eye = torch.eye(4, device=dev, dtype=torch.float32)
    for n in range(N):
        padded = torch.zeros(16, D, H, W, device=dev, dtype=torch.float32)
        padded[:C] = feat_volume[n]
        vol = ncdhw_to_vol4(padded)
        dv = _f32c(depth[n, 0, :, 0])
        w30 = volume_warp_setup(eye, _f32c(pose[n]), _f32c(cam_intr[n])).reshape(1, 30)
        key_t = torch.ones_like(vol)
        h = est_attend(key_t, [vol], [vol], w30, dv, depth_min, depth_interval, align_corners=align_corners)
        vol4_to_ncdhw(h, out[n])
    return out[:, :C]


def depthlayer(logits, depth_values):
    """Same contract as ``depthlayer`` (hybrid_depth_decoder.py:33-38): logits [B,D,H,W], depth_values [B,D,H,W]
    (constant over H,W) -> (depth [B,1,H,W], prob [B,1,H,W])."""
    B, D, H, W = logits.shape
    depth = torch.empty(B, 1, H, W, device=logits.device, dtype=torch.float32)
    prob = torch.empty_like(depth)
    for b in range(B):
        dv = _f32c(depth_values[min(b, depth_values.shape[0] - 1), :, 0, 0])
        head_softargmin(dv, logits_in=_f32c(logits[b]), depth_out=depth[b, 0], prob_out=prob[b, 0], up=1)
    return depth, prob
