"""spe_b200 — B200-native heatmap -> 6-DoF pose stage of mohsij/spacecraft-pose-estimation.

Public surface (mirrors the reference's call sites):
    get_max_preds, get_final_preds        (landmark_regression/lib/core/inference.py)
    accuracy                              (landmark_regression/lib/core/evaluate.py:42-80)
    PnPSolver.solve / solvePnPRansac      (pose_estimation/export_predicted_poses_real.py:177-204)
    boxes.pick_boxes / boxes.xywh2cs      (object_detection/export_object_detection_bounding_boxes.py:313-329,
                                           landmark_regression/lib/dataset/PEdataset.py:98-113)
    HeatmapToPose                         decode + pose without leaving the device, sharded over ranks
The arithmetic lives in libspe_b200.so (CUDA, sm_100a) behind include/spe_b200.h.
"""
from . import boxes, models, synth  # noqa: F401
from ._lib import LIB_PATH, SpeError  # noqa: F401
from .evaluate import accuracy  # noqa: F401
from .inference import decode_device, get_final_preds, get_final_preds_combined, get_max_preds  # noqa: F401
from .pnp import PnPSolver, PoseBatch, solvePnPRansac  # noqa: F401

__all__ = ["accuracy", "get_max_preds", "get_final_preds", "get_final_preds_combined", "decode_device", "PnPSolver", "PoseBatch", "solvePnPRansac", "boxes", "models", "synth", "SpeError", "LIB_PATH"]
