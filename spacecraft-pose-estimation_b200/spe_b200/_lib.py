"""ctypes binding of libspe_b200.so (include/spe_b200.h).  There is no CPU fallback: if the
library is missing or CUDA is unavailable every entry point raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_size_t, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspe_b200.so")

_lib = None
ABI_VERSION = 2


class SpeError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpeError(
                f"{LIB_PATH} not found: build it with `python spacecraft-pose-estimation_b200/build.py` "
                "(or __graft_entry__.build()). There is no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        _declare(L)
        if L.spe_abi_version() != ABI_VERSION:
            raise SpeError("libspe_b200.so ABI version mismatch")
        _lib = L
    return _lib


def _declare(L):
    fp, ip, up, dp = c_void_p, c_void_p, c_void_p, c_void_p  # device pointers travel as integers
    L.spe_abi_version.restype = c_int
    L.spe_status_string.restype = c_char_p
    L.spe_status_string.argtypes = [c_int]
    L.spe_last_cuda_error.restype = c_char_p
    L.spe_max_preds_f32.restype = c_int
    L.spe_max_preds_f32.argtypes = [fp, c_int, c_int, c_int, c_int, fp, fp, ip, c_void_p]
    L.spe_decode_f32.restype = c_int
    L.spe_decode_f32.argtypes = [fp, c_int, c_int, c_int, c_int, fp, fp, c_int, fp, fp, ip, c_void_p]
    L.spe_decode_kpts_f32.restype = c_int
    L.spe_decode_kpts_f32.argtypes = [fp, c_int, c_int, c_int, c_int, fp, fp, c_int, fp, ip, c_void_p]
    L.spe_decode_kpts_ex_f32.restype = c_int
    L.spe_decode_kpts_ex_f32.argtypes = [fp, c_int, c_int, c_int, c_int, fp, fp, c_int, fp, ip, c_int, c_void_p]
    L.spe_decode_combined_kpts_f32.restype = c_int
    L.spe_decode_combined_kpts_f32.argtypes = [POINTER(c_void_p), c_int, c_int, ip, c_int, c_int, c_int, c_int, c_int, fp, fp, c_int, fp, ip,
                                               c_void_p]
    L.spe_boxes_to_center_scale_f64.restype = c_int
    L.spe_boxes_to_center_scale_f64.argtypes = [dp, c_int, fp, fp, c_void_p]
    L.spe_pick_boxes_f32.restype = c_int
    L.spe_pick_boxes_f32.argtypes = [fp, fp, ip, c_int, c_int, c_double, c_double, dp, fp, ip, fp, fp, c_void_p]
    L.spe_pck_counts_f32.restype = c_int
    L.spe_pck_counts_f32.argtypes = [fp, fp, c_int, c_int, c_double, c_double, c_double, ip, c_void_p]
    if hasattr(L, "spe_pnp_model_create"):
        L.spe_pnp_model_create.restype = c_int
        L.spe_pnp_model_create.argtypes = [POINTER(c_double), c_int, POINTER(c_double), POINTER(c_double), c_int, POINTER(c_void_p)]
        L.spe_pnp_model_destroy.restype = c_int
        L.spe_pnp_model_destroy.argtypes = [c_void_p]
        L.spe_pnp_model_num_landmarks.restype = c_int
        L.spe_pnp_model_num_landmarks.argtypes = [c_void_p]
        L.spe_pnp_model_minimal_sets.restype = c_int
        L.spe_pnp_model_minimal_sets.argtypes = [c_void_p, c_int, c_int, POINTER(c_int32)]
        L.spe_pnp_control_entry.restype = c_int
        L.spe_pnp_control_entry.argtypes = [POINTER(c_double), c_int, POINTER(c_int32), POINTER(c_float), POINTER(ctypes.c_int64)]
        L.spe_ransac_workspace_bytes.restype = c_size_t
        L.spe_ransac_workspace_bytes.argtypes = [c_void_p, c_int, c_int]
        L.spe_ransac_epnp_f32.restype = c_int
        L.spe_ransac_epnp_f32.argtypes = [c_void_p, fp, c_int, c_int, c_float, c_double, c_float, fp, up, ip, ip, dp,
                                          c_void_p, c_size_t, c_int, c_void_p]
        L.spe_ransac_score_f32.restype = c_int
        L.spe_ransac_score_f32.argtypes = [c_void_p, fp, c_int, c_int, c_float, c_double, c_float, c_void_p, c_size_t, c_int, c_void_p]
        L.spe_ransac_select_refit_f32.restype = c_int
        L.spe_ransac_select_refit_f32.argtypes = [c_void_p, c_int, c_int, c_double, fp, up, ip, ip, dp, c_void_p, c_size_t, c_int, c_void_p]
        L.spe_pnp_minimal_sets_host.restype = c_int
        L.spe_pnp_minimal_sets_host.argtypes = [c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]
        L.spe_ransac_replay_f64.restype = c_int
        L.spe_ransac_replay_f64.argtypes = [c_void_p, c_int, c_int, c_float, c_double, c_void_p, c_size_t, c_void_p]
        L.spe_ransac_read_budget.restype = c_int
        L.spe_ransac_read_budget.argtypes = [c_void_p, c_void_p, c_int, c_int, ip, c_void_p]
        L.spe_ransac_debug_scores.restype = c_int
        L.spe_ransac_debug_scores.argtypes = [c_void_p, c_void_p, c_int, c_int, ip, up, c_void_p]
        L.spe_pipeline_workspace_bytes.restype = c_size_t
        L.spe_pipeline_workspace_bytes.argtypes = [c_void_p, c_int, c_int, c_int]
        L.spe_heatmap_to_pose_f32.restype = c_int
        L.spe_heatmap_to_pose_f32.argtypes = [c_void_p, fp, c_int, c_int, c_int, c_int, fp, fp, c_int, c_int, c_float, c_double,
                                              c_float, fp, up, ip, fp, c_void_p, c_size_t, c_int, c_void_p]


FLAG_REFINE_LM = 1
FLAG_ADAPTIVE = 2
FLAG_BACKGROUND_TAIL = 4
FLAG_JACOBI_SVD = 8
FLAG_EXACT = 16
MAX_HYPOTHESES = 16384
DECODE_BACKGROUND = 1

EXPORTED_SYMBOLS = (
    "spe_abi_version", "spe_status_string", "spe_last_cuda_error", "spe_max_preds_f32", "spe_decode_f32",
    "spe_decode_kpts_f32", "spe_decode_kpts_ex_f32", "spe_decode_combined_kpts_f32", "spe_boxes_to_center_scale_f64", "spe_pick_boxes_f32", "spe_pck_counts_f32", "spe_pnp_model_create", "spe_pnp_model_destroy", "spe_pnp_model_num_landmarks",
    "spe_pnp_model_minimal_sets", "spe_pnp_minimal_sets_host", "spe_pnp_control_entry", "spe_ransac_replay_f64", "spe_ransac_read_budget", "spe_ransac_workspace_bytes", "spe_ransac_epnp_f32", "spe_ransac_score_f32",
    "spe_ransac_select_refit_f32", "spe_ransac_debug_scores",
    "spe_pipeline_workspace_bytes", "spe_heatmap_to_pose_f32",
)


def check(status: int, what: str):
    if status != 0:
        L = lib()
        msg = L.spe_status_string(status).decode()
        if status == -2:
            msg += ": " + L.spe_last_cuda_error().decode()
        raise SpeError(f"{what} failed: {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise SpeError("spe_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch
