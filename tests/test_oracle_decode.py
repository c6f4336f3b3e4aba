"""The decode oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py ran core.inference / utils.transforms from /root/reference)."""
import numpy as np
import pytest

from oracle import decode_ref


def _cases(g):
    return [str(n) for n in g["names"]]


def test_get_max_preds_matches_reference(decode_golden):
    g = decode_golden
    for name in _cases(g):
        preds, maxvals, idx = decode_ref.get_max_preds(g[f"{name}/hm"], return_index=True)
        assert preds.dtype == np.float32 and maxvals.dtype == np.float32
        np.testing.assert_array_equal(preds, g[f"{name}/max_preds"], err_msg=name)
        np.testing.assert_array_equal(maxvals, g[f"{name}/maxvals"], err_msg=name)
        np.testing.assert_array_equal(idx, g[f"{name}/argmax"], err_msg=name)


@pytest.mark.parametrize("fn", [decode_ref.get_final_preds, decode_ref.get_final_preds_fast])
def test_get_final_preds_matches_reference(decode_golden, fn):
    g = decode_golden
    for name in _cases(g):
        hm, c, s = g[f"{name}/hm"], g[f"{name}/center"], g[f"{name}/scale"]
        hm_before = hm.copy()
        preds, maxvals = fn(True, hm, c, s)
        np.testing.assert_array_equal(preds, g[f"{name}/final_preds"], err_msg=name)
        np.testing.assert_array_equal(maxvals, g[f"{name}/maxvals"], err_msg=name)
        preds_nopp, _ = fn(False, hm, c, s)
        np.testing.assert_array_equal(preds_nopp, g[f"{name}/final_preds_nopp"], err_msg=name)
        np.testing.assert_array_equal(hm, hm_before)  # inputs are not mutated


def test_transform_preds_known_answers(decode_golden):
    g = decode_golden
    out = decode_ref.transform_preds(g["tp/coords"], g["tp/center"], g["tp/scale"], [64, 64])
    np.testing.assert_array_equal(out, g["tp/out64"])
    # SURVEY App. E.1 (typed in from the survey, independent of the npz)
    np.testing.assert_allclose(out, [[749.796875, 491.265625], [650.5, 290.25], [1260.8125, 900.5625]], rtol=0, atol=1e-9)
    t = decode_ref.get_affine_transform(np.array([960, 600], np.float32), np.array([3, 2.5], np.float32), 0, [64, 64], inv=1)
    np.testing.assert_allclose(t, [[9.375, 0, 660], [0, 9.375, 300]], atol=1e-9)
    t = decode_ref.get_affine_transform(np.array([960, 600], np.float32), np.array([3, 2.5], np.float32), 0, [72, 96], inv=1)
    np.testing.assert_allclose(t, [[25 / 3, 0, 660], [0, 25 / 3, 200]], atol=1e-9)


def test_closed_form_affine_replays_reference(decode_golden):
    """App. A.4: the float32-replaying closed form the CUDA kernel uses is within 1e-9 px of the
    reference's transform_preds in float64 and bit-equal after float32 rounding in > 99.9 %."""
    sweep = decode_golden["tp/sweep"]
    worst, same, total = 0.0, 0, 0
    for row in sweep:
        W, H = int(row[0]), int(row[1])
        c, s = row[2:4].astype(np.float32), row[4:6].astype(np.float32)
        xy = row[6:18].reshape(6, 2)
        ref = row[18:30].reshape(6, 2)
        ax, bx, ay, by = decode_ref.inverse_affine_closed_form(c, s, W, H)
        mine = np.stack([ax * xy[:, 0] + bx, ay * xy[:, 1] + by], 1)
        worst = max(worst, np.abs(mine - ref).max())
        same += int((mine.astype(np.float32) == ref.astype(np.float32)).sum())
        total += mine.size
    assert worst < 1e-9
    assert same / total > 0.999


def test_edge_semantics():
    """App. A.1-A.3 stated directly."""
    hm = np.zeros((1, 4, 8, 8), np.float32)
    hm[0, 0] = -1.0  # all negative
    hm[0, 1, 3, 4] = np.nan
    hm[0, 1, 5, 5] = 9.0
    hm[0, 2, 2, 6] = 1.0
    hm[0, 2, 4, 1] = 1.0  # tie, later
    hm[0, 3, 4, 4] = 1.0
    hm[0, 3, 4, 5] = 0.5  # +x neighbour larger than -x neighbour
    hm[0, 3, 3, 4] = 0.5  # -y neighbour larger
    c = np.array([[100.0, 50.0]], np.float32)
    s = np.array([[0.32, 0.1]], np.float32)  # sw = 64 -> a = 8
    preds, maxvals, idx = decode_ref.get_final_preds(True, hm, c, s, return_index=True)
    assert idx[0].tolist() == [0, 3 * 8 + 4, 2 * 8 + 6, 4 * 8 + 4]
    assert maxvals[0, 0, 0] == -1.0 and np.isnan(maxvals[0, 1, 0])
    ax, bx, ay, by = decode_ref.inverse_affine_closed_form(c[0], s[0], 8, 8)
    np.testing.assert_allclose(preds[0, 0], [bx, by], atol=1e-4)  # masked -> heatmap (0,0)
    np.testing.assert_allclose(preds[0, 1], [bx, by], atol=1e-4)  # NaN max is not > 0
    np.testing.assert_allclose(preds[0, 2], [ax * 6 + bx, ay * 2 + by], atol=1e-4)  # px = W-2 refinable but flat
    np.testing.assert_allclose(preds[0, 3], [ax * 4.25 + bx, ay * 3.75 + by], atol=1e-4)


def test_asserts_like_reference():
    with pytest.raises(AssertionError):
        decode_ref.get_max_preds([[1.0]])
    with pytest.raises(AssertionError):
        decode_ref.get_max_preds(np.zeros((2, 3, 4), np.float32))
