"""detection boxes -> (center, scale).  CPU: the oracle restatement against golden vectors produced by
executing the reference's own source lines (tests/golden/make_boxes_golden.py).  GPU: spe_pick_boxes_f32 /
spe_boxes_to_center_scale_f64 through the C ABI, bit-exact against the same vectors."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "boxes_golden.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN)


def test_oracle_equals_reference_golden(g):
    from oracle import boxes_ref

    for b in range(len(g["counts"])):
        W, H = g["sizes"][g["size_id"][b]]
        n = int(g["counts"][b])
        bb, sc = boxes_ref.pick_box(g["boxes"][b, :n], g["scores"][b, :n], W, H)
        np.testing.assert_array_equal(np.array(bb, np.float64), g["xywh"][b])
        assert (np.isnan(sc) and np.isnan(g["best_score"][b])) or sc == g["best_score"][b]
        c, s = boxes_ref.xywh2cs(*bb)
        assert c.dtype == np.float32 and s.dtype == np.float32
        np.testing.assert_array_equal(c, g["center"][b])
        np.testing.assert_array_equal(s, g["scale"][b])
    for row, c0, s0 in zip(g["q_xywh"], g["q_center"], g["q_scale"]):
        c, s = boxes_ref.xywh2cs(*row)
        np.testing.assert_array_equal(c, c0)
        np.testing.assert_array_equal(s, s0)
    # the center[0] == -1 quirk of the reference: no factor 1.5
    np.testing.assert_array_equal(g["q_scale"][0], np.array([2.0 / 200, 30.0 / 200], np.float32))


@pytest.mark.gpu
def test_pick_boxes_and_xywh2cs_bit_exact_on_gpu(g):
    from spe_b200 import boxes

    for sid, (W, H) in enumerate(g["sizes"]):
        sel = g["size_id"] == sid
        out = boxes.pick_boxes(g["boxes"][sel], g["scores"][sel], g["counts"][sel], W, H)
        np.testing.assert_array_equal(out["xywh"], g["xywh"][sel])
        np.testing.assert_array_equal(out["center"], g["center"][sel])
        np.testing.assert_array_equal(out["scale"], g["scale"][sel])
        np.testing.assert_array_equal(np.isnan(out["score"]), np.isnan(g["best_score"][sel]))
        ok = ~np.isnan(out["score"])
        np.testing.assert_array_equal(out["score"][ok].astype(np.float64), g["best_score"][sel][ok])
        whole = ~np.isin(g["counts"][sel], (1, 2))
        assert np.all(out["index"][whole] == -1) and np.all(out["index"][~whole] >= 0)
    c, s = boxes.xywh2cs(g["q_xywh"])
    np.testing.assert_array_equal(c, g["q_center"])
    np.testing.assert_array_equal(s, g["q_scale"])


@pytest.mark.gpu
def test_boxes_feed_the_decode_without_a_host_hop(g):
    """torch in -> torch out: the (center, scale) tensors go straight into get_final_preds on the device, and an empty
    detector output (K = 0) takes the whole image everywhere."""
    import torch

    import spe_b200
    from oracle import boxes_ref, decode_ref
    from spe_b200 import boxes

    sel = g["size_id"] == 0
    bt, st, ct = (torch.from_numpy(np.ascontiguousarray(g[k][sel])).cuda() for k in ("boxes", "scores", "counts"))
    out = boxes.pick_boxes(bt, st, ct, 1920, 1200)
    assert out["center"].is_cuda and torch.equal(out["center"].cpu(), torch.from_numpy(g["center"][sel]))
    B = int(sel.sum())
    hm = torch.rand((B, 3, 16, 16), device="cuda")
    preds, maxvals = spe_b200.get_final_preds(True, hm, out["center"], out["scale"])
    rp, rm = decode_ref.get_final_preds_fast(True, hm.cpu().numpy(), g["center"][sel], g["scale"][sel])
    assert np.abs(preds.cpu().numpy() - rp).max() <= 1.3e-4
    empty = boxes.pick_boxes(np.zeros((5, 0, 4), np.float32), np.zeros((5, 0), np.float32), None, 640, 480)
    c, s = boxes_ref.xywh2cs(0, 0, 640, 480)
    assert np.all(empty["center"] == c) and np.all(empty["scale"] == s) and np.all(empty["index"] == -1)
