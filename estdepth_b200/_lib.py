"""ctypes binding of libestdepth_b200.so (the C ABI declared in include/estdepth_b200.h).

There is no fallback: if the shared library has not been built (``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C estdepth_b200/csrc``) every op raises.  The library is kept in-tree, next to this file.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libestdepth_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_void_p = ctypes.c_void_p


class ConvDesc(ctypes.Structure):
    """struct estd_conv3d_desc (include/estdepth_b200.h)."""
    _fields_ = [
        ("in0", c_void_p), ("in0_chunks", ctypes.c_int),
        ("in1", c_void_p), ("in1_chunks", ctypes.c_int),
        ("weight", c_void_p), ("weight_tc", c_void_p), ("precision", ctypes.c_int), ("status", c_void_p),
        ("planar", ctypes.c_int), ("dilation", ctypes.c_int),
        ("scale", c_void_p), ("shift", c_void_p),
        ("cout_pad", ctypes.c_int), ("act_split", ctypes.c_int), ("act_lo", ctypes.c_int), ("act_hi", ctypes.c_int),
        ("res0", c_void_p), ("res1", c_void_p),
        ("post_scale", ctypes.c_float),
        ("out0", c_void_p), ("out0_chunks", ctypes.c_int),
        ("out1", c_void_p), ("out1_chunks", ctypes.c_int),
        ("gn_partials", c_void_p),
        ("D", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
        ("in0_split", ctypes.c_int), ("in1_split", ctypes.c_int), ("res_split", ctypes.c_int), ("out_split", ctypes.c_int),
        ("head_w", c_void_p), ("head_b", c_void_p), ("head_out", c_void_p),
        ("out_up2", ctypes.c_int),
    ]


# name -> (restype, argtypes); must list every symbol the header declares (tests/test_abi.py checks it)
_I, _F, _P = ctypes.c_int, ctypes.c_float, c_void_p
SIGNATURES = {
    "estd_version": (_I, []),
    "estd_last_error": (ctypes.c_char_p, []),
    "estd_launch_count": (ctypes.c_ulonglong, []),
    "estd_homography_setup": (_I, [_P, _P, _P, _P, _P]),
    "estd_homography_table": (_I, [_P, _I, _P, ctypes.POINTER(ctypes.c_int), _I, _P, _P]),
    "estd_homography_from_proj": (_I, [_P, _P, _P, _P]),
    "estd_volume_warp_setup": (_I, [_P, _P, _P, _P, _P]),
    "estd_volume_warp_table": (_I, [ctypes.POINTER(_P), _I, _P, ctypes.POINTER(ctypes.c_int), _I, _P, _P]),
    "estd_premix": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "estd_premix_batch": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "estd_warp_cost": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "estd_conv3d_num_ctas": (_I, [ctypes.POINTER(ConvDesc)]),
    "estd_conv3d": (_I, [ctypes.POINTER(ConvDesc), _P]),
    "estd_est_attend": (_I, [_P, _I, ctypes.POINTER(_P), ctypes.POINTER(_P), _P, _P, _F, _F, _P, _I, _I, _I, _I, _P]),
    "estd_gn_finalize": (_I, [_P, _I, _I, ctypes.c_double, _F, _P, _P]),
    "estd_gru_reset": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "estd_gru_blend": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "estd_head_softargmin": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "estd_vol4_to_ncdhw": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "estd_ncdhw_to_vol4": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "estd_scalar_to_vol4": (_I, [_P, _P, _I, _I, _I, _P]),
    "estd_nchw_to_vol4": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "estd_stem_conv": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "estd_stem7_conv": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "estd_maxpool3x3s2_vol4": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "estd_upsample_bilinear_vol4": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "estd_vol4_to_nchw": (_I, [_P, _P, _I, _I, _I, _I, _P]),
}

_lib = None


def get():
    """Returns the loaded library; raises RuntimeError (loudly) when it is missing or incomplete."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "estdepth_b200: %s not found -- build it with `make -C estdepth_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU/PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if lib.estd_version() < 100:
            raise RuntimeError("estdepth_b200: stale library (version %d)" % lib.estd_version())
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = get().estd_last_error()
        raise RuntimeError("estdepth_b200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count():
    return int(get().estd_launch_count())
