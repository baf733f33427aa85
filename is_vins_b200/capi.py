"""ctypes binding of the C ABI in ``include/isv_capi.h`` (libisv_b200.so).

The shared library is the product; this module only loads it and mirrors its PODs.  There is no
CPU fallback: if the library is missing, or no sm_100 GPU is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISV_B200_LIB", os.path.join(_HERE, "libisv_b200.so"))

ISV_OK, ISV_ERR_BAD_ARG, ISV_ERR_CUDA, ISV_ERR_ALLOC = 0, 1, 2, 3
W_NOT_SPD, W_RANK_DEFICIENT, W_NONFINITE, W_NONUNIT_QUAT, W_EIG_NOCONV, W_SINGULAR = 1, 2, 4, 8, 16, 32
W_BAD_INDEX = 64
W_DIAG_COUPLED = 128
IN_PTS_I_Z_ONE, IN_TRI_RECORDS, OUT_TRI_RECORDS = 1, 2, 4   # isv_batch_in::flags (the TRI bits: host-pointer entry point only)
SE3_TRI_REC, REL_TRI_REC, VB_TRI_REC, RP_IN_TRI_REC, PG_TRI_REC, RP_TRI_REC = 33, 33, 54, 4, 59, 12
IMU_JAC_REC, YAW_REC = 480, 4
ACC_REC, ACC_COVREL, ACC_DISTANCE, ACC_LENGTH, ACC_VIO_INDEX, ACC_PG_INDEX, ACC_TS = 119, 48, 84, 85, 86, 87, 88
ACC_RI, ACC_TI, ACC_RP_VALID, ACC_RP, ACC_COVABS = 89, 98, 101, 102, 115
RUN_FORWARD, RUN_BACKWARD, RUN_BOTH = 1, 2, 3
RUN_FORWARD_STAGE1, RUN_FORWARD_STAGE2, RUN_FACTOR_JAC, RUN_BACKWARD_STAGE2 = 4, 8, 16, 32
TUNE_FUSED_MAX_WINDOWS, TUNE_EVENT_MODE, TUNE_ACC_PERSIST = 1, 2, 3   # isv_set_tuning knobs (include/isv_capi.h)
POSE, SB, SE3_REC, REL_REC, VB_REC, RP_IN_REC, RP_REC, PG_REC, PREINT_REC = 7, 9, 48, 48, 90, 5, 13, 89, 467
IMU_RAW_REC = 7

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class isv_config(C.Structure):
    _fields_ = [("alpha", C.c_double), ("proj_sqrt_info", C.c_double * 4), ("g", C.c_double * 3),
                ("acc_n", C.c_double), ("gyr_n", C.c_double), ("acc_w", C.c_double), ("gyr_w", C.c_double),
                ("vo_size", C.c_int), ("all_buf_size", C.c_int), ("qr_rank_eps_log10", C.c_int),
                ("reserved", C.c_int)]


class isv_batch_in(C.Structure):
    _fields_ = [("n_windows", C.c_int32), ("ex_pose_shared", C.c_int32), ("lm_offset", C.c_void_p),
                ("lm_obs", C.c_void_p), ("lm_stride", C.c_int64), ("pose_fwd", C.c_void_p),
                ("ex_pose", C.c_void_p), ("prior_se3", C.c_void_p), ("prior_rel", C.c_void_p),
                ("prior_rp", C.c_void_p), ("pose_bwd", C.c_void_p), ("sb_bwd", C.c_void_p),
                ("prior_vb", C.c_void_p), ("preint", C.c_void_p),
                # ABI 2: raw IMU samples instead of the pre-integration record, and the ISV_IN_* flags
                ("imu_raw", C.c_void_p), ("imu_init", C.c_void_p), ("imu_count", C.c_void_p),
                ("imu_k_max", C.c_int32), ("flags", C.c_int32),
                # ABI 3: pts_i.x / pts_i.y as the FP32 values the feature tracker produced
                ("lm_xy_f32", C.c_void_p)]


class isv_batch_out(C.Structure):
    _fields_ = [("se3_out", C.c_void_p), ("pg_out", C.c_void_p), ("rel_out", C.c_void_p),
                ("vb_out", C.c_void_p), ("rp_out", C.c_void_p), ("rank", C.c_void_p), ("status", C.c_void_p)]


class isv_forensic_out(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("lamda_prior_fwd", "g_bwd", "kld_fwd", "kld_bwd", "lamda_prior_bwd", "eig_bwd",
                                          "info_abs", "info_yaw")]


class isv_init_in(C.Structure):
    _fields_ = [("n_windows", C.c_int32), ("poses", C.c_void_p), ("sbs", C.c_void_p), ("preint", C.c_void_p)]


class isv_init_out(C.Structure):
    _fields_ = [("rel_out", C.c_void_p), ("se3_out", C.c_void_p), ("vb_out", C.c_void_p), ("rank", C.c_void_p),
                ("status", C.c_void_p)]


class isv_preint_in(C.Structure):
    _fields_ = [("n", C.c_int32), ("k_max", C.c_int32), ("k_count", C.c_void_p), ("imu_raw", C.c_void_p),
                ("imu_init", C.c_void_p)]


class isv_fwd_in(C.Structure):
    _fields_ = [("n_landmarks", C.c_int32), ("pose0", c_double_p), ("pose1", c_double_p), ("ex_pose", c_double_p),
                ("inv_dep", c_double_p), ("pts_i", c_double_p), ("pts_j", c_double_p), ("prior_se3", c_double_p),
                ("prior_rel", c_double_p), ("prior_rp", c_double_p)]


class isv_fwd_out(C.Structure):
    _fields_ = [("se3", C.c_double * SE3_REC), ("pg", C.c_double * PG_REC), ("rank", C.c_int32),
                ("status", C.c_int32)]


class isv_bwd_in(C.Structure):
    _fields_ = [("pose_i", c_double_p), ("sb_i", c_double_p), ("pose_j", c_double_p), ("sb_j", c_double_p),
                ("prior_vb", c_double_p), ("preint", c_double_p)]


class isv_bwd_out(C.Structure):
    _fields_ = [("rel", C.c_double * REL_REC), ("vb", C.c_double * VB_REC), ("rp", C.c_double * RP_REC),
                ("rank", C.c_int32), ("status", C.c_int32)]


class isv_param_blocks(C.Structure):
    _fields_ = [("n_pose", C.c_int32), ("n_speed_bias", C.c_int32), ("n_ex_pose", C.c_int32),
                ("n_feature", C.c_int32), ("pose", C.c_void_p), ("speed_bias", C.c_void_p),
                ("ex_pose", C.c_void_p), ("feature", C.c_void_p)]


class isv_proj_factors(C.Structure):
    _fields_ = [("n", C.c_int64), ("stride", C.c_int64), ("idx", C.c_void_p), ("obs", C.c_void_p),
                ("cauchy_a", C.c_double), ("td_obs", C.c_void_p), ("td", C.c_void_p), ("td_idx", C.c_void_p),
                ("n_td", C.c_int32), ("reserved", C.c_int32), ("tr_over_row", C.c_double)]


class isv_proj_eval(C.Structure):
    _fields_ = [("residuals", C.c_void_p), ("jac_pose_i", C.c_void_p), ("jac_pose_j", C.c_void_p),
                ("jac_ex_pose", C.c_void_p), ("jac_feature", C.c_void_p), ("jac_td", C.c_void_p)]


class isv_imu_factors(C.Structure):
    _fields_ = [("n", C.c_int32), ("idx", C.c_void_p), ("preint", C.c_void_p)]


class isv_imu_eval(C.Structure):
    _fields_ = [("residuals", C.c_void_p), ("jacobians", C.c_void_p)]


class isv_small_factors(C.Structure):
    _fields_ = [("n_rel", C.c_int32), ("n_se3", C.c_int32), ("n_vb", C.c_int32), ("n_rp", C.c_int32),
                ("n_yaw", C.c_int32), ("rel_idx", C.c_void_p), ("rel_rec", C.c_void_p), ("se3_idx", C.c_void_p),
                ("se3_rec", C.c_void_p), ("vb_idx", C.c_void_p), ("vb_rec", C.c_void_p), ("rp_idx", C.c_void_p),
                ("rp_rec", C.c_void_p), ("yaw_idx", C.c_void_p), ("yaw_rec", C.c_void_p), ("cauchy_a", C.c_double)]


class isv_small_eval(C.Structure):
    _fields_ = [("rel_res", C.c_void_p), ("rel_jac", C.c_void_p), ("se3_res", C.c_void_p), ("se3_jac", C.c_void_p),
                ("vb_res", C.c_void_p), ("vb_jac", C.c_void_p), ("rp_res", C.c_void_p), ("rp_jac", C.c_void_p),
                ("yaw_res", C.c_void_p), ("yaw_jac", C.c_void_p)]


class isv_seq_update_in(C.Structure):
    _fields_ = [("old_P", C.c_void_p), ("old_R", C.c_void_p), ("old_vb", C.c_void_p), ("pose", C.c_void_p),
                ("speed_bias", C.c_void_p)]


class isv_seq_frame(C.Structure):
    _fields_ = [("ex_pose_shared", C.c_int32), ("lm_offset", C.c_void_p), ("lm_obs", C.c_void_p),
                ("lm_stride", C.c_int64), ("pose_fwd", C.c_void_p), ("ex_pose", C.c_void_p),
                ("pose_bwd", C.c_void_p), ("sb_bwd", C.c_void_p), ("preint", C.c_void_p), ("ts", C.c_void_p),
                ("Ri", C.c_void_p), ("ti", C.c_void_p)]


class isv_seq_host(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("rel", "se3", "vb", "rp", "rp_valid", "acc", "pg_count", "last_se3",
                                          "last_pg", "last_rel", "last_vb", "last_rp", "last_rank", "last_status")]


class isv_prior_block(C.Structure):
    _fields_ = [("global_size", C.c_int32), ("idx", C.c_int32), ("x_offset", C.c_int32), ("pos", C.c_int32)]


class isv_marg_prior(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_blocks", C.c_int32), ("blocks", C.c_void_p), ("linearized_jacobians", C.c_void_p),
                ("linearized_residuals", C.c_void_p), ("x0", C.c_void_p), ("x", C.c_void_p)]


# every symbol include/isv_capi.h declares: (name, restype, argtypes)
_H = C.c_void_p
SYMBOLS = [
    ("isv_default_config", None, [C.POINTER(isv_config)]),
    ("isv_abi_version", C.c_int, []),
    ("isv_status_string", C.c_char_p, [C.c_int]),
    ("isv_create", C.c_int, [C.POINTER(isv_config), C.c_int, C.POINTER(_H)]),
    ("isv_destroy", None, [_H]),
    ("isv_stream", C.c_void_p, [_H]),
    ("isv_set_stream", C.c_int, [_H, C.c_void_p]),
    ("isv_synchronize", C.c_int, [_H]),
    ("isv_launch_count", C.c_int64, [_H]),
    ("isv_set_tuning", C.c_int, [_H, C.c_int, C.c_int]),
    ("isv_order_map_init", C.c_int, [C.c_int, c_int32_p]),
    ("isv_order_map_forward", C.c_int, [C.c_int, c_int32_p]),
    ("isv_order_map_backward", C.c_int, [C.c_int, c_int32_p]),
    ("isv_marg_window_batch", C.c_int, [_H, C.POINTER(isv_batch_in), C.POINTER(isv_batch_out), C.c_int]),
    ("isv_marg_window_batch_host", C.c_int, [_H, C.POINTER(isv_batch_in), C.POINTER(isv_batch_out), C.c_int]),
    ("isv_marg_forensic_batch", C.c_int, [_H, C.POINTER(isv_batch_in), C.POINTER(isv_batch_out), C.POINTER(isv_forensic_out)]),
    ("isv_literal_schur", C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("isv_marg_forward", C.c_int, [_H, C.POINTER(isv_fwd_in), C.POINTER(isv_fwd_out)]),
    ("isv_marg_backward", C.c_int, [_H, C.POINTER(isv_bwd_in), C.POINTER(isv_bwd_out)]),
    ("isv_marg_event", C.c_int, [_H, C.POINTER(isv_fwd_in), C.POINTER(isv_bwd_in), C.POINTER(isv_fwd_out), C.POINTER(isv_bwd_out)]),
    ("isv_test_fast_special", C.c_int, [_H, C.c_int, c_double_p, c_double_p, c_double_p]),
    ("isv_test_fused_stamps", C.c_int, [_H, C.POINTER(isv_batch_in), C.POINTER(isv_batch_out), C.c_void_p]),
    ("isv_test_event_latency", C.c_int, [_H, C.POINTER(isv_fwd_in), C.POINTER(isv_bwd_in), C.POINTER(isv_fwd_out),
                                         C.POINTER(isv_bwd_out), C.c_int, C.POINTER(C.c_double)]),
    ("isv_init_sparsify_batch", C.c_int, [_H, C.POINTER(isv_init_in), C.POINTER(isv_init_out)]),
    ("isv_init_sparsify_host", C.c_int, [_H, C.POINTER(isv_init_in), C.POINTER(isv_init_out)]),
    ("isv_preintegrate_batch", C.c_int, [_H, C.POINTER(isv_preint_in), C.c_void_p]),
    ("isv_preintegrate_host", C.c_int, [_H, C.POINTER(isv_preint_in), C.c_void_p]),
    ("isv_eval_projection_batch", C.c_int, [_H, C.POINTER(isv_param_blocks), C.POINTER(isv_proj_factors),
                                            C.POINTER(isv_proj_eval), C.c_void_p]),
    ("isv_eval_imu_batch", C.c_int, [_H, C.POINTER(isv_param_blocks), C.POINTER(isv_imu_factors),
                                     C.POINTER(isv_imu_eval), C.c_void_p]),
    ("isv_eval_small_batch", C.c_int, [_H, C.POINTER(isv_param_blocks), C.POINTER(isv_small_factors),
                                       C.POINTER(isv_small_eval), C.c_void_p]),
    ("isv_eval_problem", C.c_int, [_H, C.POINTER(isv_param_blocks), C.POINTER(isv_proj_factors),
                                   C.POINTER(isv_proj_eval), C.POINTER(isv_imu_factors), C.POINTER(isv_imu_eval),
                                   C.POINTER(isv_small_factors), C.POINTER(isv_small_eval), C.c_void_p]),
    ("isv_seq_create", C.c_int, [_H, C.c_int, C.POINTER(C.c_void_p)]),
    ("isv_seq_destroy", None, [_H, C.c_void_p]),
    ("isv_seq_init", C.c_int, [_H, C.c_void_p, C.POINTER(isv_init_in), C.c_void_p]),
    ("isv_seq_update", C.c_int, [_H, C.c_void_p, C.POINTER(isv_seq_update_in)]),
    ("isv_seq_yaw", C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("isv_seq_marginalize", C.c_int, [_H, C.c_void_p, C.POINTER(isv_seq_frame), C.c_double, C.c_void_p, C.c_void_p]),
    ("isv_seq_export_host", C.c_int, [_H, C.c_void_p, C.POINTER(isv_seq_host)]),
    ("isv_seq_import_host", C.c_int, [_H, C.c_void_p, C.POINTER(isv_seq_host)]),
    ("isv_build_normal_equations", C.c_int, [_H, C.c_void_p, C.c_void_p]),
    ("isv_marginalize_generic", C.c_int, [_H, C.c_void_p, C.c_void_p]),
    ("isv_reduced_system", C.c_int, [_H, C.c_void_p, C.c_void_p]),
    ("isv_marginalize_host", C.c_int, [_H, C.c_void_p, C.c_void_p]),
    ("isv_eval_marg_prior", C.c_int, [_H, C.POINTER(isv_marg_prior), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("isv_add_marg_prior", C.c_int, [_H, C.POINTER(isv_marg_prior), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    ("isv_schur_eig", C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32]),
    ("isv_pose_plus_batch", C.c_int, [_H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("isv_pose_plus_jacobian", None, [c_double_p]),
    ("isv_test_psd_eig", C.c_int, [_H, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_int32_p]),
    ("isv_test_sym_eig", C.c_int, [_H, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_int32_p]),
]

_lib: Optional[C.CDLL] = None


class IsvError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen libisv_b200.so (built by ``__graft_entry__.build()`` / ``make -C is_vins_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IsvError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "-- there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)     # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != ISV_OK:
        name = load().isv_status_string(status).decode()
        raise IsvError(f"{what or 'isv call'} failed: {name}")


def default_config() -> isv_config:
    cfg = isv_config()
    load().isv_default_config(C.byref(cfg))
    return cfg
