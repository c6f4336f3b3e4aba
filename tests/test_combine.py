"""SURVEY §8 f2 — heatmap pre-combination (flip test, model ensemble) fused into the decode.
CPU: the oracle restatement against the reference's own flip_back (golden) and torch.
GPU: the fused kernel against oracle-combine + oracle-decode, and against the reference's exact
torch expressions evaluated on the device followed by the plain decode kernel."""
import numpy as np
import pytest

from oracle import combine_ref, decode_ref


def _maps(rng, B, J, H, W, k):
    out = []
    for _ in range(k):
        hm = rng.normal(scale=0.05, size=(B, J, H, W)).astype(np.float32)
        peaks = rng.integers(0, H * W, (B, J))
        flat = hm.reshape(B, J, -1)
        np.put_along_axis(flat, peaks[..., None], rng.uniform(0.5, 1.0, (B, J, 1)).astype(np.float32), axis=2)
        out.append(hm)
    return out


def test_oracle_flip_back_and_means_cpu():
    import torch

    rng = np.random.default_rng(0)
    a, f = _maps(rng, 2, 6, 8, 12, 2)
    pairs = [(0, 1), (3, 5)]
    fb = combine_ref.flip_back(f, pairs)
    # definition: reversed along W, matched joints swapped
    np.testing.assert_array_equal(fb[:, 0], f[:, 1, :, ::-1])
    np.testing.assert_array_equal(fb[:, 2], f[:, 2, :, ::-1])
    np.testing.assert_array_equal(fb[:, 5], f[:, 3, :, ::-1])
    assert combine_ref.flip_perm(6, pairs).tolist() == [1, 0, 2, 5, 4, 3]
    # the reference's torch expressions (function.py:356-366) on CPU tensors
    of = torch.from_numpy(fb.copy())
    of[:, :, :, 1:] = of.clone()[:, :, :, 0:-1]
    ref = ((torch.from_numpy(a) + of) * 0.5).numpy()
    np.testing.assert_array_equal(combine_ref.flip_average(a, f, pairs, shift_heatmap=True), ref)
    ref_ns = ((torch.from_numpy(a) + torch.from_numpy(fb.copy())) * 0.5).numpy()
    np.testing.assert_array_equal(combine_ref.flip_average(a, f, pairs, shift_heatmap=False), ref_ns)
    for k in (1, 2, 4):  # exact reciprocals: CPU true division == CUDA reciprocal multiply
        outs = _maps(rng, 1, 3, 8, 8, k)
        t = torch.from_numpy(outs[0].copy())
        for o in outs[1:]:
            t += torch.from_numpy(o)
        np.testing.assert_array_equal(combine_ref.ensemble_mean(outs), (t / k).numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(16, 11, 64, 64), (4, 17, 96, 72), (3, 5, 17, 19), (2, 4, 128, 128)])
def test_fused_flip_average_gpu(shape):
    import torch

    import spe_b200

    B, J, H, W = shape
    rng = np.random.default_rng(B + J + H)
    a, f = _maps(rng, B, J, H, W, 2)
    a[0, 0, H // 2, W // 2] = np.nan
    c = np.stack([rng.uniform(100, 1800, B), rng.uniform(100, 1100, B)], 1).astype(np.float32)
    s = rng.uniform(0.3, 9.0, (B, 2)).astype(np.float32)
    pairs = [(0, 1), (2, 4)] if J >= 5 else []
    for shift in (True, False):
        comb = combine_ref.flip_average(a, f, pairs, shift_heatmap=shift)
        rp, rm, ri = decode_ref.get_final_preds_fast(True, comb, c, s, return_index=True)
        p, m, idx = spe_b200.get_final_preds_combined(True, [a, f], c, s, mode="flip", flip_pairs=pairs, shift_heatmap=shift, return_index=True)
        np.testing.assert_array_equal(idx, ri.astype(np.int32))
        np.testing.assert_array_equal(m, rm)
        assert np.abs(p.view(np.int32).astype(np.int64) - rp.view(np.int32).astype(np.int64)).max() <= 1
    # the reference's own sequence on the device: torch ops, then the plain decode
    ta, tf = torch.from_numpy(a).cuda(), torch.from_numpy(f).cuda()
    of = torch.from_numpy(combine_ref.flip_back(tf.cpu().numpy(), pairs).copy()).cuda()
    of[:, :, :, 1:] = of.clone()[:, :, :, 0:-1]
    out = (ta + of) * 0.5
    p2, m2, i2 = spe_b200.get_final_preds(True, out, torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda(), return_index=True)
    p1, m1, i1 = spe_b200.get_final_preds_combined(True, [ta, tf], torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda(), mode="flip",
                                                   flip_pairs=pairs, shift_heatmap=True, return_index=True)
    assert torch.equal(i1, i2) and torch.equal(p1, p2)
    assert torch.equal(torch.nan_to_num(m1, nan=-7.0), torch.nan_to_num(m2, nan=-7.0))


@pytest.mark.gpu
@pytest.mark.parametrize("K", [1, 2, 3, 6])
def test_fused_ensemble_mean_gpu(K):
    import torch

    import spe_b200

    B, J, H, W = 8, 11, 64, 64
    rng = np.random.default_rng(K)
    outs = _maps(rng, B, J, H, W, K)
    c = np.stack([rng.uniform(100, 1800, B), rng.uniform(100, 1100, B)], 1).astype(np.float32)
    s = rng.uniform(0.3, 9.0, (B, 2)).astype(np.float32)
    comb = combine_ref.ensemble_mean(outs)
    rp, rm, ri = decode_ref.get_final_preds_fast(True, comb, c, s, return_index=True)
    p, m, idx = spe_b200.get_final_preds_combined(True, outs, c, s, mode="mean", return_index=True)
    np.testing.assert_array_equal(idx, ri.astype(np.int32))
    np.testing.assert_array_equal(m, rm)
    assert np.abs(p.view(np.int32).astype(np.int64) - rp.view(np.int32).astype(np.int64)).max() <= 1
    # validate_cv's expressions on CUDA tensors (function.py:531-536), then the plain decode
    touts = [torch.from_numpy(o).cuda() for o in outs]
    output = touts[0].clone()
    for o in touts[1:]:
        output += o
    output = output / len(touts)
    p2, m2, i2 = spe_b200.get_final_preds(True, output, torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda(), return_index=True)
    assert np.array_equal(idx, i2.cpu().numpy()) and np.array_equal(m, m2.cpu().numpy()) and np.array_equal(p, p2.cpu().numpy())
