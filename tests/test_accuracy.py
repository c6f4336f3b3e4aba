"""accuracy() (landmark_regression/lib/core/evaluate.py:42-80): the oracle against golden vectors made
by the imported reference (CPU), and the device path against both (GPU, through the C ABI)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "accuracy_golden.npz"))


def _case(g, name):
    src = name[:-len("_thr02")] if name.endswith("_thr02") else name
    return g[f"{src}/output"], g[f"{src}/target"], float(g[f"{name}/thr"])


def test_oracle_accuracy_equals_the_reference(golden):
    from oracle import accuracy_ref

    for name in golden["cases"]:
        output, target, thr = _case(golden, name)
        acc, avg_acc, cnt, pred = accuracy_ref.accuracy(output, target, "gaussian", thr)
        np.testing.assert_array_equal(acc, golden[f"{name}/acc"])
        assert avg_acc == float(golden[f"{name}/avg_acc"]) and cnt == int(golden[f"{name}/cnt"])
        np.testing.assert_array_equal(pred, golden[f"{name}/pred"])
        assert pred.dtype == np.float32 and acc.dtype == np.float64
    # the cases contain what they were built for: joints that are never counted, an all-invalid batch, a changed `thr`
    assert golden["square32/acc"][1] == -1 and int(golden["all_invalid/cnt"]) == 0
    np.testing.assert_array_equal(golden["square32/acc"], golden["square32_thr02/acc"])


@pytest.mark.gpu
def test_device_accuracy_is_bit_identical_to_the_reference(golden):
    import torch

    import spe_b200

    for name in golden["cases"]:
        output, target, thr = _case(golden, name)
        acc, avg_acc, cnt, pred = spe_b200.accuracy(output, target, "gaussian", thr)  # NumPy in -> NumPy out
        np.testing.assert_array_equal(acc, golden[f"{name}/acc"])
        assert avg_acc == float(golden[f"{name}/avg_acc"]) and cnt == int(golden[f"{name}/cnt"])
        np.testing.assert_array_equal(pred, golden[f"{name}/pred"])
        assert isinstance(pred, np.ndarray) and pred.dtype == np.float32
        acc_t, avg_t, cnt_t, pred_t = spe_b200.accuracy(torch.from_numpy(output).cuda(), torch.from_numpy(target).cuda())  # CUDA in -> CUDA pred
        np.testing.assert_array_equal(acc_t, acc)
        assert avg_t == avg_acc and cnt_t == cnt and pred_t.is_cuda
        np.testing.assert_array_equal(pred_t.cpu().numpy(), pred)


@pytest.mark.gpu
def test_device_accuracy_against_the_oracle_on_a_training_sized_batch():
    """256 x 11 x 64 x 64, peaks within a few pixels of the targets so that distances fall on both sides of 0.5."""
    import torch

    import spe_b200
    from oracle import accuracy_ref

    rng = np.random.default_rng(5)
    B, J, H, W = 256, 11, 64, 64
    tgt_xy = np.stack([rng.integers(0, W, (B, J)), rng.integers(0, H, (B, J))], -1)
    prd_xy = np.clip(tgt_xy + rng.integers(-4, 5, (B, J, 2)), 0, [W - 1, H - 1])
    output = rng.normal(scale=0.01, size=(B, J, H, W)).astype(np.float32)
    target = np.zeros((B, J, H, W), np.float32)
    bi, ji = np.meshgrid(np.arange(B), np.arange(J), indexing="ij")
    output[bi, ji, prd_xy[..., 1], prd_xy[..., 0]] = 1.0
    target[bi, ji, tgt_xy[..., 1], tgt_xy[..., 0]] = 1.0
    ref = accuracy_ref.accuracy(output, target)
    got = spe_b200.accuracy(torch.from_numpy(output).cuda(), torch.from_numpy(target).cuda())
    np.testing.assert_array_equal(got[0], ref[0])
    assert got[1] == ref[1] and got[2] == ref[2] and 0.2 < ref[1] < 0.9
    np.testing.assert_array_equal(got[3].cpu().numpy(), ref[3])
    with pytest.raises(ValueError):
        spe_b200.accuracy(output, target, hm_type="offset")
