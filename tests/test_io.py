"""DepthMapWriter writes the files the reference's drivers write (eval_hybrid.py:260-264,276-277,282-286,306-307)."""
import os

import numpy as np
import pytest
import torch

from estdepth_b200.io import DepthMapWriter


def _reference_bytes(tmp_path, name, array):
    path = os.path.join(tmp_path, name)
    np.save(path, array)
    with open(path, "rb") as f:
        return f.read()


def test_writer_files_equal_the_drivers_formula(tmp_path):
    g = torch.Generator().manual_seed(0)
    depth = torch.rand(1, 1, 48, 64, generator=g) * 10            # outputs[("depth", t, s)]
    prob = torch.rand(1, 1, 48, 64, generator=g)                  # outputs[("init_prob", t)]
    depth[0, 0, 0, :4] = torch.tensor([0.1, 65504.0, 1e-8, 2049.0])        # fp16 corner cases: exact, max, underflow, tie
    with DepthMapWriter(max_pending=2) as w:
        for i in range(5):                                        # more maps than the queue holds: back-pressure, order kept
            w.save(depth + i, os.path.join(tmp_path, "d%d.npy" % i))
        w.save(prob.squeeze(), os.path.join(tmp_path, "p.npy"), squeeze_channel=False)
        w.save_outputs({("depth", 0, 0): depth}, {("depth", 0, 0): os.path.join(tmp_path, "o.npy")})
    for i in range(5):
        want = _reference_bytes(tmp_path, "ref_d%d.npy" % i, np.float16((depth + i).squeeze(1).cpu().numpy()))   # eval_hybrid.py:260
        with open(os.path.join(tmp_path, "d%d.npy" % i), "rb") as f:
            assert f.read() == want
    want = _reference_bytes(tmp_path, "ref_p.npy", np.float16(prob.squeeze().cpu().numpy()))                       # eval_hybrid.py:276-277
    with open(os.path.join(tmp_path, "p.npy"), "rb") as f:
        assert f.read() == want
    got = np.load(os.path.join(tmp_path, "o.npy"))
    assert got.dtype == np.float16 and got.shape == (1, 48, 64)


def test_writer_reports_write_errors(tmp_path):
    w = DepthMapWriter()
    w.save(torch.zeros(1, 1, 4, 4), os.path.join(tmp_path, "missing_dir", "x.npy"))
    with pytest.raises(OSError):
        w.close()
