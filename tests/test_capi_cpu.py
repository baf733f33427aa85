"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/isv_capi.h
declares, the index maps are bit exact, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from is_vins_b200 import capi
from oracle import isv_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(isv_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    declared = _declared_symbols()
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libisv_b200.so does not export {name}"
    bound = {s[0] for s in capi.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    assert lib.isv_abi_version() == 4


def test_default_config_matches_euroc_yaml():
    cfg = capi.default_config()
    ocfg = O.Config()
    assert cfg.alpha == ocfg.alpha and cfg.vo_size == ocfg.vo_size and cfg.all_buf_size == ocfg.all_buf_size
    assert list(cfg.proj_sqrt_info) == [460.0, 0.0, 0.0, 460.0]
    assert list(cfg.g) == [0.0, 0.0, ocfg.g_norm]
    assert (cfg.acc_n, cfg.gyr_n, cfg.acc_w, cfg.gyr_w) == (ocfg.acc_n, ocfg.gyr_n, ocfg.acc_w, ocfg.gyr_w)
    assert cfg.qr_rank_eps_log10 == -16


def test_order_maps_match_oracle_bit_exact():
    lib = capi.load()
    for V in (2, 8, 12):
        buf = (C.c_int32 * (4 * V))()
        nb = lib.isv_order_map_init(V, buf)
        om = O.order_map_init(V)
        keys = [("pose", i) for i in range(V)] + [("sb", V - 1)] + [("sb", i) for i in range(V - 1)]
        assert nb == 2 * V
        assert [(buf[2 * i], buf[2 * i + 1]) for i in range(nb)] == [om[k] for k in keys]
    for L in (0, 1, 150, 1000):
        buf = (C.c_int32 * (2 * (L + 2)))()
        nb = lib.isv_order_map_forward(L, buf)
        om = O.order_map_forward(L)
        keys = [("pose", 1), ("pose", 0)] + [("feat", k) for k in range(L)]
        assert nb == L + 2
        assert [(buf[2 * i], buf[2 * i + 1]) for i in range(nb)] == [om[k] for k in keys]
    buf = (C.c_int32 * 8)()
    assert lib.isv_order_map_backward(8, buf) == 4
    om = O.order_map_backward(8)
    assert [(buf[2 * i], buf[2 * i + 1]) for i in range(4)] == [om[("pose", 8)], om[("sb", 8)], om[("pose", 7)],
                                                                 om[("sb", 7)]]


def test_bad_arguments_and_no_gpu_fail_loudly():
    import torch
    lib = capi.load()
    cfg = capi.default_config()
    h = C.c_void_p()
    assert lib.isv_create(None, 0, C.byref(h)) == capi.ISV_ERR_BAD_ARG
    bad = capi.default_config()
    bad.vo_size = 1
    assert lib.isv_create(C.byref(bad), 0, C.byref(h)) == capi.ISV_ERR_BAD_ARG
    assert lib.isv_status_string(capi.ISV_ERR_CUDA) == b"ISV_ERR_CUDA"
    if not torch.cuda.is_available():
        # no device: the product refuses to run (there is no CPU fallback to fall into)
        assert lib.isv_create(C.byref(cfg), 0, C.byref(h)) == capi.ISV_ERR_CUDA
        from is_vins_b200 import MargBackend
        with pytest.raises(capi.IsvError):
            MargBackend(0)
    assert lib.isv_marg_window_batch(None, None, None, 3) == capi.ISV_ERR_BAD_ARG
    assert lib.isv_launch_count(None) == 0


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "is_vins_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("oracle/ (test", "").replace("`oracle/`", ""), fn


def test_pose_local_parameterization_jacobian_is_identity_over_zero():
    """PoseLocalParameterization::ComputeJacobian (src/factor/pose_local_parameterization.cpp:20-27): [I6 ; 0], 7 x 6
    row-major -- host-side constant, the same bytes as the oracle's."""
    lib = capi.load()
    j = np.full(42, np.nan)
    lib.isv_pose_plus_jacobian(j.ctypes.data_as(capi.c_double_p))
    assert np.array_equal(j.reshape(7, 6), O.pose_compute_jacobian())
