"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/spe_b200.h declares; argument errors are reported without touching a GPU; the
product path refuses to run without CUDA instead of falling back to the oracle."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry

    from spe_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        entry.build()
    return _lib.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spe_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from spe_b200 import _lib

    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in spe_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_abi_basics_without_gpu(lib):
    assert lib.spe_abi_version() == 1
    assert lib.spe_status_string(0) == b"ok"
    assert b"invalid" in lib.spe_status_string(-1)
    # argument validation happens before any CUDA call
    assert lib.spe_decode_f32(None, -1, 11, 64, 64, None, None, 1, None, None, None, None) == -1
    assert lib.spe_decode_f32(None, 0, 11, 64, 64, None, None, 1, None, None, None, None) == 0  # empty batch
    assert lib.spe_max_preds_f32(None, 4, 11, 64, 64, None, None, None, None) == -1  # null heatmaps
    assert lib.spe_max_preds_f32(None, 4, 0, 64, 64, None, None, None, None) == -1
    handle = ctypes.c_void_p()
    lm = (ctypes.c_double * 9)(*([0.0] * 9))
    K = (ctypes.c_double * 9)(*([1.0] * 9))
    assert lib.spe_pnp_model_create(lm, 3, K, None, 64, ctypes.byref(handle)) == -1  # J < 4
    assert lib.spe_pnp_model_create(lm, 33, K, None, 64, ctypes.byref(handle)) == -1  # J > 32
    assert lib.spe_pnp_model_create(lm, 11, K, None, 0, ctypes.byref(handle)) == -1
    assert lib.spe_pnp_model_destroy(None) == 0
    assert lib.spe_ransac_workspace_bytes(None, 4, 64) == 0


def test_product_refuses_to_run_without_cuda():
    import torch

    import spe_b200

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(spe_b200.SpeError):
        spe_b200.get_max_preds(np.zeros((1, 2, 8, 8), np.float32))
    with pytest.raises(spe_b200.SpeError):
        spe_b200.PnPSolver(spe_b200.models.tango().landmarks, spe_b200.models.tango().K)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spacecraft-pose-estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+cv2\b", src, flags=re.M), f"{f} must not depend on OpenCV"


def test_pose_entry_points_validate_arguments_without_gpu(lib):
    """No model can exist without a GPU, so every pose entry point must reject a NULL model (and
    the combined decode its argument errors) before touching CUDA."""
    assert lib.spe_ransac_epnp_f32(None, None, 4, 64, 15.0, 0.99, -1.0, None, None, None, None, None, None, 0, 0, None) == -1
    assert lib.spe_ransac_score_f32(None, None, 4, 64, 15.0, 0.99, -1.0, None, 0, 0, None) == -1
    assert lib.spe_ransac_select_refit_f32(None, 4, 64, 0.99, None, None, None, None, None, None, 0, 0, None) == -1
    assert lib.spe_heatmap_to_pose_f32(None, None, 4, 11, 64, 64, None, None, 1, 64, 15.0, 0.99, -1.0, None, None, None, None, None, 0, 0, None) == -1
    assert lib.spe_pipeline_workspace_bytes(None, 4, 11, 64) == 0
    assert lib.spe_pnp_model_num_landmarks(None) == -1
    ptrs = (ctypes.c_void_p * 2)(None, None)
    assert lib.spe_decode_combined_kpts_f32(ptrs, 0, 0, None, 0, 4, 11, 64, 64, None, None, 1, None, None, None) == -1  # K < 1
    assert lib.spe_decode_combined_kpts_f32(ptrs, 9, 0, None, 0, 4, 11, 64, 64, None, None, 1, None, None, None) == -1  # K > 8
    assert lib.spe_decode_combined_kpts_f32(ptrs, 3, 1, None, 0, 4, 11, 64, 64, None, None, 1, None, None, None) == -1  # flip needs K == 2
    assert lib.spe_decode_combined_kpts_f32(ptrs, 2, 7, None, 0, 4, 11, 64, 64, None, None, 1, None, None, None) == -1  # unknown mode
    assert lib.spe_decode_combined_kpts_f32(ptrs, 2, 0, None, 0, 0, 11, 64, 64, None, None, 1, None, None, None) == 0  # empty batch
