"""Host logic of ``DepthNetHybrid`` on CPU: the whole prepare/fuse flow with the CUDA library replaced by a recorder.

Every ``estd_*`` entry point is a no-op that returns success, so tensors hold garbage -- what is checked is everything the
host decides: which kernels are enqueued how often (the launch schedule DESIGN.md section 4b describes), the shapes /
keys / ordering of what a driver gets back, and the hidden-state protocol with its quirks (Q3 first window
without EST, Q4 stale memory pose, the ESTM memory FIFO of eval_hybrid_seq.py:102-116).  The numerics of the same flow are
the GPU tests' job (tests/test_gpu_model.py)."""
import collections
import ctypes

import pytest
import torch

from estdepth_b200 import _lib, model as model_mod, ops, sharding, synth
from tests.helpers import synth_model_and_state


class RecorderLib(object):
    def __init__(self):
        self.calls = collections.Counter()
        self.attend_sources = []             # n_src of every estd_est_attend call

    def __getattr__(self, name):
        if not name.startswith("estd_"):
            raise AttributeError(name)

        def entry(*args):
            self.calls[name] += 1
            if name == "estd_est_attend":
                self.attend_sources.append(args[1])
            return 148 if name == "estd_conv3d_num_ctas" else 0
        return entry


@pytest.fixture()
def recorder(monkeypatch):
    lib = RecorderLib()
    monkeypatch.setattr(_lib, "get", lambda: lib)
    monkeypatch.setattr(ops, "_ptr", lambda t, dtype=torch.float32: None if t is None else ctypes.c_void_p(t.data_ptr()))
    monkeypatch.setattr(ops, "_act_ptr", lambda t, dtype=torch.float32: None if t is None else t.data_ptr())
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "status_flag", lambda device: torch.zeros(1, dtype=torch.int32))
    monkeypatch.setattr(model_mod, "_upload", lambda t, dev: t.contiguous())
    return lib


def _model(**kw):
    m, _ = synth_model_and_state(18, 32)
    m.overlap_context = 0                    # single stream: there is no CUDA stream to fork on the CPU
    for k, v in kw.items():
        setattr(m, k, v)
    return m


def _window(start, views=5):
    imgs, poses, K, sample = synth.synth_inputs(views, 128, 160, seed=0, start=start)
    return imgs, poses, K


@pytest.mark.parametrize("merged", [True, False])
def test_joint_windows_launch_schedule_outputs_and_stale_pose(recorder, merged):
    m = _model(merged_pre2=merged)
    with torch.no_grad():
        imgs, poses, K = _window(0)
        out1, state1, pose1 = m._forward_val(imgs, poses, K, None, None)
        c1 = collections.Counter(recorder.calls)
        recorder.calls.clear()
        imgs2, poses2, K2 = _window(3)
        out2, state2, pose2 = m._forward_val(imgs2, poses2, K2, state1, pose1)
        c2 = collections.Counter(recorder.calls)
    pre_convs = 3 if merged else 4
    # window 1 (quirk Q3: no EST): per target pre1/pre2 + dres0 x2 + dres1 x2 + dres2 + value|key + head0 + head1
    assert c1["estd_conv3d"] == 3 * (pre_convs + 8) and c1["estd_est_attend"] == 0 and c1["estd_gru_blend"] == 0
    # window 2: + gate conv and output conv of the ConvGRU per target; one attention gather per target
    assert c2["estd_conv3d"] == 3 * (pre_convs + 10) and c2["estd_est_attend"] == 3
    assert recorder.attend_sources[:3] == [3, 3, 3]      # each target attends to the 2 other targets + 1 memory volume
    assert c2["estd_gn_finalize"] == 6 and c2["estd_gru_reset"] == 3 and c2["estd_gru_blend"] == 3
    for c in (c1, c2):
        assert c["estd_premix_batch"] == 1 and c["estd_warp_cost"] == 6 and c["estd_head_softargmin"] == 6
        assert c["estd_vol4_to_ncdhw"] == 2 and c["estd_scalar_to_vol4"] == 3
    # what a driver gets back (hybrid_depth_decoder.py:208-209,260,279,290; model_hybrid.py:183-184)
    for out in (out1, out2):
        assert set(out) == {("depth", t, s) for t in range(3) for s in range(4)} | {("init_prob", t) for t in range(3)} | \
            {("fused_prob", t) for t in range(3)}
        assert all(tuple(v.shape) == (1, 1, 128, 160) and v.dtype == torch.float32 for v in out.values())
    for state, pose in ((state1, pose1), (state2, pose2)):
        assert list(state) == ["keys", "values"] and len(state["keys"]) == len(state["values"]) == len(pose) == 1
        assert tuple(state["keys"][0].shape) == tuple(state["values"][0].shape) == (1, 16, 32, 32, 40)
        assert tuple(pose[0].shape) == (1, 4, 4)
    assert torch.equal(pose1[0], poses[:, 3])            # last target of window 1
    assert torch.equal(pose2[0], pose1[0])               # quirk Q4: the memory's pose again, not window 2's last target
    m.fix_stale_pose = True
    with torch.no_grad():
        _, _, pose2_fixed = m._forward_val(imgs2, poses2, K2, state1, pose1)
    assert torch.equal(pose2_fixed[0], poses2[:, 3])


def test_estm_protocol_memory_fifo_and_source_counts(recorder):
    """3-frame windows, memory of the last two states (eval_hybrid_seq.py:169-193): 0, 1, 2, 2 memory volumes attended."""
    m = _model()
    mem = []
    attends = []
    with torch.no_grad():
        for step in range(4):
            imgs, poses, K = _window(step, views=3)
            pre = sharding._flatten_memory(mem)
            recorder.calls.clear()
            out, costs, cposes = m._forward_val(imgs, poses, K, pre[0], pre[1])
            attends.append(recorder.calls["estd_est_attend"])
            assert recorder.calls["estd_conv3d"] == (11 if step == 0 else 13)
            assert set(k[1] for k in out) == {0} and len(costs["keys"]) == 1
            mem.append((costs, cposes))
            if len(mem) > 2:
                mem.pop(0)
    assert attends == [0, 1, 1, 1]                        # one target per window: one fused gather ...
    assert recorder.attend_sources == [1, 2, 2]           # ... over 1 / 2 / 2 memory volumes (the FIFO keeps two)


def test_batch_of_sequences_is_a_loop_of_single_sequence_pipelines(recorder):
    """Quirk Q16: the reference only runs B = 1; B > 1 here = the same schedule per batch element."""
    m = _model()
    a, b = _window(0), _window(3)
    imgs, poses, K = (torch.cat([x, y]) for x, y in zip(a, b))
    with torch.no_grad():
        out, state, pose = m._forward_val(imgs, poses, K, None, None)
    assert recorder.calls["estd_conv3d"] == 2 * 33 and recorder.calls["estd_premix_batch"] == 2
    assert all(tuple(v.shape) == (2, 1, 128, 160) for v in out.values())
    assert tuple(state["keys"][0].shape) == (2, 16, 32, 32, 40) and tuple(pose[0].shape) == (2, 4, 4)


def test_val_mode_only_and_cpu_tensors_rejected():
    m = _model()
    imgs, poses, K = _window(0)
    with pytest.raises(NotImplementedError):
        m(imgs, poses, K, None, mode="train")
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(imgs, poses, K, None, mode="val")
    with pytest.raises(AssertionError):
        with torch.no_grad():
            m.prepare(imgs[:, :2], poses[:, :2], K)      # views_num must exceed 2 (model_hybrid.py:123)


def test_frame_ids_reach_the_feature_cache_through_the_public_flow(recorder):
    """ESTM windows share 2 of their 3 frames: with frame ids only the new frame goes through the matching-feature net."""
    m = _model()
    batches = []
    inner = m.matchingFeature.forward
    m.matchingFeature.forward = lambda x: (batches.append(x.shape[0]), inner(x))[1]
    mem = []
    with torch.no_grad():
        for step in range(3):
            imgs, poses, K = _window(step, views=3)
            pre = sharding._flatten_memory(mem)
            _, costs, cposes = m._forward_val(imgs, poses, K, pre[0], pre[1], frame_ids=[step, step + 1, step + 2])
            mem = (mem + [(costs, cposes)])[-2:]
    assert batches == [3, 1, 1]
    batches.clear()
    with torch.no_grad():
        m._forward_val(*_window(0, views=3), None, None)
    assert batches == [3]                                 # no ids: the reference's behaviour, every view recomputed
