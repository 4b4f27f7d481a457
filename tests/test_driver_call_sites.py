"""The reference's own eval drivers against the drop-in boundary, statically: every call expression `DepthNetHybrid(...)` and
`model(...)` in eval_hybrid.py / eval_hybrid_seq.py must bind to this repository's constructor / forward signatures, and every
`outputs[(...)]` key the drivers read must be a key `forward` returns.  The drivers themselves need a GPU and the
datasets; this parses their source (read-only, under /root/reference) and is skipped where the reference tree is absent."""
import ast
import inspect
import os

import pytest

from estdepth_b200 import DepthNetHybrid

REF = "/root/reference"
DRIVERS = ["eval_hybrid.py", "eval_hybrid_seq.py"]
pytestmark = pytest.mark.skipif(not all(os.path.exists(os.path.join(REF, d)) for d in DRIVERS), reason="reference tree not present")


def _calls(tree, name):
    return [n for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id == name]


@pytest.mark.parametrize("driver", DRIVERS)
def test_driver_call_expressions_bind_to_the_drop_in_signatures(driver):
    tree = ast.parse(open(os.path.join(REF, driver)).read())
    ctor_sig = inspect.signature(DepthNetHybrid.__init__)
    fwd_sig = inspect.signature(DepthNetHybrid.forward)
    ctors, forwards = _calls(tree, "DepthNetHybrid"), _calls(tree, "model")
    assert ctors and forwards, "the driver constructs the model and calls it"
    for call, sig in [(c, ctor_sig) for c in ctors] + [(c, fwd_sig) for c in forwards]:
        assert not any(isinstance(a, ast.Starred) for a in call.args) and all(k.arg is not None for k in call.keywords)
        bound = sig.bind(None, *([None] * len(call.args)), **{k.arg: None for k in call.keywords})      # raises TypeError on a mismatch
        if sig is fwd_sig:
            # the drivers pass imgs, cam_poses, cam_intr, sample positionally and the memory + mode by keyword
            assert list(bound.arguments)[1:5] == ["imgs", "cam_poses", "cam_intr", "sample"]
            mode = [k.value for k in call.keywords if k.arg == "mode"]
            assert mode and isinstance(mode[0], ast.Constant) and mode[0].value == "val"                  # the only mode the hot path implements
            assert {"pre_costs", "pre_cam_poses"} <= set(bound.arguments)
    # the call unpacks three results
    unpack = [n for n in ast.walk(tree) if isinstance(n, ast.Assign) and isinstance(n.value, ast.Call) and n.value in forwards]
    assert unpack and all(isinstance(n.targets[0], ast.Tuple) and len(n.targets[0].elts) == 3 for n in unpack)


@pytest.mark.parametrize("driver", DRIVERS)
def test_driver_reads_only_output_keys_forward_returns(driver):
    tree = ast.parse(open(os.path.join(REF, driver)).read())
    kinds = set()
    for n in ast.walk(tree):
        if isinstance(n, ast.Subscript) and isinstance(n.value, ast.Name) and n.value.id == "outputs" and isinstance(n.slice, ast.Tuple):
            head = n.slice.elts[0]
            assert isinstance(head, ast.Constant)
            scale = n.slice.elts[2].value if len(n.slice.elts) == 3 and isinstance(n.slice.elts[2], ast.Constant) else None
            kinds.add((head.value, len(n.slice.elts), scale))
    assert kinds, "the driver reads the outputs dict"
    for kind, arity, scale in kinds:
        assert (kind, arity) in {("depth", 3), ("init_prob", 2), ("fused_prob", 2)}      # model.py: ("depth", t, s), ("init_prob", t), ("fused_prob", t)
        assert scale is None or scale in (0, 1, 2, 3)
