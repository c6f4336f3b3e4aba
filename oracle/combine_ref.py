"""Oracle (test infrastructure): the heatmap pre-combinations the reference applies before
decoding, restated in NumPy float32.

  flip_back            landmark_regression/lib/utils/transforms.py:15-29
  flip-test average    landmark_regression/lib/core/function.py:347-366
  model-ensemble mean  landmark_regression/lib/core/function.py:525-536 (validate_cv)

The ensemble mean is evaluated by torch on the GPU in the reference (`output/len(models)` on CUDA
tensors), where tensor / python-scalar is a multiplication by the float32 reciprocal
(ATen BinaryDivTrueKernel: "compute a * reciprocal(b)"); the restatement follows that, and the GPU
parity test additionally checks the fused kernel against the same torch expression run on the device.
"""
from __future__ import annotations

import numpy as np


def flip_back(output_flipped: np.ndarray, matched_parts) -> np.ndarray:
    assert output_flipped.ndim == 4, "output_flipped should be [batch_size, num_joints, height, width]"
    out = output_flipped[:, :, :, ::-1].copy()
    for a, b in matched_parts:
        tmp = out[:, a].copy()
        out[:, a] = out[:, b]
        out[:, b] = tmp
    return out


def flip_average(output: np.ndarray, output_flipped_raw: np.ndarray, matched_parts=(), shift_heatmap: bool = False) -> np.ndarray:
    """function.py:347-366 — output_flipped_raw is the network's output on the mirrored input."""
    f = flip_back(np.asarray(output_flipped_raw, np.float32), matched_parts)
    if shift_heatmap:  # "feature is not aligned, shift flipped heatmap for higher accuracy"
        g = f.copy()
        g[:, :, :, 1:] = f[:, :, :, 0:-1]
        f = g
    return ((np.asarray(output, np.float32) + f) * np.float32(0.5)).astype(np.float32)


def ensemble_mean(outputs) -> np.ndarray:
    """function.py:525-536 — sequential float32 sum, then * float32(1/K) (torch CUDA semantics)."""
    acc = np.asarray(outputs[0], np.float32).copy()
    for o in outputs[1:]:
        acc += np.asarray(o, np.float32)
    return (acc * np.float32(np.float32(1.0) / np.float32(len(outputs)))).astype(np.float32)


def flip_perm(num_joints: int, matched_parts) -> np.ndarray:
    """perm[j] = joint of the flipped tensor whose map lands on joint j after flip_back."""
    perm = np.arange(num_joints, dtype=np.int32)
    for a, b in matched_parts:
        perm[a], perm[b] = perm[b], perm[a]
    return perm
