"""GPU parity of the decode kernels against (1) golden vectors produced by the reference itself
and (2) the oracle on seeded inputs, through the C ABI.  Bars: argmax indices bit-exact,
coordinates <= 1e-4 px (north_star); in practice bit-equal except rare 1-ulp cases above 1024 px
(SURVEY App. A.4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_PX = 1e-4


def _spe():
    import spe_b200

    return spe_b200


def _ulp_ok(a, b):
    """equal, or one float32 ulp apart (the documented residue of cv2.getAffineTransform's LU)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    both_nan = np.isnan(a) & np.isnan(b)
    one_ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)) <= 1
    return both_nan | one_ulp


def test_golden_from_reference(decode_golden):
    spe = _spe()
    g = decode_golden
    exact, total = 0, 0
    for name in [str(n) for n in g["names"]]:
        hm, c, s = g[f"{name}/hm"], g[f"{name}/center"], g[f"{name}/scale"]
        p, m, idx = spe.get_max_preds(hm, return_index=True)
        assert p.dtype == np.float32 and p.shape == g[f"{name}/max_preds"].shape and m.shape == g[f"{name}/maxvals"].shape
        np.testing.assert_array_equal(idx, g[f"{name}/argmax"].astype(np.int32), err_msg=name)
        np.testing.assert_array_equal(p, g[f"{name}/max_preds"], err_msg=name)
        np.testing.assert_array_equal(m, g[f"{name}/maxvals"], err_msg=name)
        for pp, key in ((True, "final_preds"), (False, "final_preds_nopp")):
            p, m, idx = spe.get_final_preds(pp, hm, c, s, return_index=True)
            ref = g[f"{name}/{key}"]
            np.testing.assert_array_equal(idx, g[f"{name}/argmax"].astype(np.int32), err_msg=name)
            np.testing.assert_array_equal(m, g[f"{name}/maxvals"], err_msg=name)
            assert _ulp_ok(p, ref).all(), name
            exact += int((p == ref).sum())
            total += p.size
            small = np.abs(ref) < 1024
            assert np.abs(p - ref)[small].max(initial=0) <= TOL_PX, name
    assert exact / total > 0.995, (exact, total)


@pytest.mark.parametrize("shape", [(64, 11, 64, 64), (16, 17, 96, 72), (8, 11, 128, 128), (3, 5, 384, 384), (5, 3, 17, 19),
                                   (2, 4, 7, 5), (130, 2, 16, 16), (1, 1, 768, 768)])
def test_against_oracle_random(shape):
    from oracle import decode_ref

    spe = _spe()
    B, J, H, W = shape
    rng = np.random.default_rng(B * 1000 + J * 100 + H + W)
    hm = rng.normal(scale=0.3, size=shape).astype(np.float32)
    # plant ties, NaNs, infinities and all-negative maps
    flat = hm.reshape(B * J, -1)
    for k in range(0, B * J, 3):
        flat[k, rng.integers(0, H * W, 3)] = 4.0
    for k in range(1, B * J, 7):
        flat[k, rng.integers(0, H * W, 2)] = np.nan
    for k in range(2, B * J, 11):
        flat[k] = -np.abs(flat[k])
    for k in range(5, B * J, 13):
        flat[k, rng.integers(0, H * W)] = np.inf
    for k in range(6, B * J, 17):
        flat[k] = -np.inf
    c = np.stack([rng.uniform(0, 1920, B), rng.uniform(0, 1200, B)], 1).astype(np.float32)
    s = rng.uniform(0.15, 14.0, (B, 2)).astype(np.float32)
    ref_p, ref_m, ref_i = decode_ref.get_final_preds_fast(True, hm, c, s, return_index=True)
    p, m, idx = spe.get_final_preds(True, hm, c, s, return_index=True)
    np.testing.assert_array_equal(idx, ref_i.astype(np.int32))
    np.testing.assert_array_equal(m, ref_m)
    assert _ulp_ok(p, ref_p).all()
    assert (p == ref_p).mean() > 0.995
    ref_p0, ref_m0 = decode_ref.get_max_preds(hm)
    p0, m0 = spe.get_max_preds(hm)
    np.testing.assert_array_equal(p0, ref_p0)
    np.testing.assert_array_equal(m0, ref_m0)


def test_synthetic_tango_batch_and_torch_inputs():
    import torch

    from oracle import decode_ref

    spe = _spe()
    fr = spe.synth.make_frames(spe.models.tango(), 256, 64, 64, seed=spe.synth.BASE_SEED + 1)
    ref_p, ref_m, ref_i = decode_ref.get_final_preds_fast(True, fr.heatmaps, fr.center, fr.scale, return_index=True)

    class Cfg:
        class TEST:
            POST_PROCESS = True

    hm_t = torch.from_numpy(fr.heatmaps).cuda()
    hm_before = hm_t.clone()
    p, m, idx = spe.get_final_preds(Cfg, hm_t, torch.from_numpy(fr.center).cuda(), torch.from_numpy(fr.scale).cuda(), return_index=True)
    assert p.is_cuda and p.shape == (256, 11, 2) and m.shape == (256, 11, 1)
    assert torch.equal(hm_t, hm_before)  # inputs untouched
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_i.astype(np.int32))
    np.testing.assert_array_equal(m.cpu().numpy(), ref_m)
    pn = p.cpu().numpy()
    assert _ulp_ok(pn, ref_p).all() and (pn == ref_p).mean() > 0.999
    assert np.abs(pn - ref_p).max() <= 1.3e-4  # 1 ulp at > 1024 px
    # kpts layout (pred.mat rows)
    kpts, _ = spe.decode_device(hm_t, torch.from_numpy(fr.center).cuda(), torch.from_numpy(fr.scale).cuda(), True, kpts_layout=True)
    np.testing.assert_array_equal(kpts.cpu().numpy(), np.concatenate([pn, m.cpu().numpy()], -1))


def test_full_size_properties():
    """BASELINE config B size (4096 x 11 x 64 x 64): properties that need no oracle pass."""
    import torch

    spe = _spe()
    B, J, H, W = 4096, 11, 64, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    hm = torch.randn((B, J, H, W), generator=g, device="cuda") * 0.1
    peak = torch.randint(0, H * W, (B, J), generator=g, device="cuda")
    hm.view(B, J, -1).scatter_(2, peak[..., None], 3.0)
    p, m, idx = spe.get_max_preds(hm, return_index=True)
    assert torch.equal(idx.long(), peak)
    assert torch.equal(idx.long(), hm.view(B, J, -1).argmax(2))
    assert torch.all(m == 3.0)
    assert torch.equal(p[..., 0], (peak % W).float()) and torch.equal(p[..., 1], (peak // W).float())
    # idempotence / determinism
    p2, m2, idx2 = spe.get_max_preds(hm, return_index=True)
    assert torch.equal(idx, idx2) and torch.equal(p, p2)


def test_argument_errors():
    import torch

    spe = _spe()
    with pytest.raises(AssertionError):
        spe.get_max_preds(np.zeros((2, 3, 4), np.float32))
    with pytest.raises(AssertionError):
        spe.get_max_preds([[1.0]])
    with pytest.raises(ValueError):
        spe.get_final_preds(True, np.zeros((2, 3, 8, 8), np.float32), np.zeros((1, 2), np.float32), np.zeros((2, 2), np.float32))
    p, m = spe.get_max_preds(torch.zeros((0, 3, 8, 8), device="cuda"))
    assert p.shape == (0, 3, 2) and m.shape == (0, 3, 1)


def test_decode_on_many_concurrent_streams_and_after_an_empty_launch():
    """The dynamically scheduled kernel draws its maps from a claim counter.  The counter is a per-(device, stream) slot
    zeroed on the stream before every launch: 40 streams with launches in flight at once (more than the 16-entry ring of
    round 1), each launching repeatedly, must all decode their own maps completely."""
    import torch

    import spe_b200

    g = torch.Generator(device="cuda").manual_seed(11)
    streams = [torch.cuda.Stream() for _ in range(40)]
    inputs = [torch.randn((48 + 3 * i, 5, 64, 64), generator=g, device="cuda") for i in range(len(streams))]
    torch.cuda.synchronize()
    results = [[] for _ in streams]
    for rep in range(3):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                p, m, idx = spe_b200.get_max_preds(inputs[i], return_index=True)
                results[i].append((p, m, idx))
    torch.cuda.synchronize()
    for i, hm in enumerate(inputs):
        ref = hm.view(hm.shape[0], 5, -1).argmax(2)
        for p, m, idx in results[i]:
            assert torch.equal(idx.long(), ref)
            assert torch.equal(m[..., 0], hm.view(hm.shape[0], 5, -1).amax(2))
