"""The whole stage (decode -> pose) on the GPU against the oracle, on the other BASELINE configs,
plus size-independent properties at full batch sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL_DEG = 1e-3
T_TOL_REL = 1e-4


def _stage_vs_oracle(model, fr, hypotheses, min_agree):
    import cv2

    from oracle import decode_ref, pnp_ref
    from spe_b200.pipeline import HeatmapToPose

    stage = HeatmapToPose(model, hypotheses=hypotheses)
    out = stage(fr.heatmaps, fr.center, fr.scale)  # host in -> host out
    rp, rm = decode_ref.get_final_preds_fast(True, fr.heatmaps, fr.center, fr.scale)
    np.testing.assert_array_equal(out.kpts[..., 2:], rm)
    assert np.abs(out.kpts[..., :2] - rp).max() <= 1.3e-4
    ref_kpts = np.concatenate([rp, rm], -1)
    agree, n_ok = 0, 0
    for b in range(fr.heatmaps.shape[0]):
        good = pnp_ref.confidence_filter(ref_kpts[b, :, 2])
        if good.sum() < 6:
            continue
        ok, p7, mask, rv, tv = pnp_ref.pose_from_keypoints(ref_kpts[b], model.landmarks, model.K, model.dist, iterations=10000)
        assert ok == (int(out.status[b]) == 0), b
        if not ok:
            continue
        n_ok += 1
        if (int(out.inlier_mask[b]) & 0xFFFFFFFF) == mask:
            agree += 1
            rot, tr = pnp_ref.pose_errors(out.pose7[b].astype(np.float64), p7)
            assert rot <= 2e-3 and tr <= T_TOL_REL, (b, rot, tr)  # pose7 is float32: ~1e-3 deg resolution near pi
    assert n_ok > 0 and agree / n_ok >= min_agree, (agree, n_ok)
    return agree, n_ok


def test_config_c_hubble_17_landmarks_96x72():
    import spe_b200

    model = spe_b200.models.hubble_synthetic(17)
    fr = spe_b200.synth.make_frames(model, 96, 96, 72, seed=spe_b200.synth.BASE_SEED + 2, z_range=(3.0, 8.0))
    agree, n_ok = _stage_vs_oracle(model, fr, hypotheses=512, min_agree=0.95)
    print(f"config C (J=17, 96x72, H=512): winner-mask agreement {agree}/{n_ok}")


def test_config_c_hubble_24_landmarks_adaptive_filter():
    import spe_b200

    model = spe_b200.models.hubble_synthetic(24)
    fr = spe_b200.synth.make_frames(model, 64, 96, 72, seed=spe_b200.synth.BASE_SEED + 22, z_range=(3.0, 8.0))
    agree, n_ok = _stage_vs_oracle(model, fr, hypotheses=512, min_agree=0.95)
    print(f"config C (J=24, 96x72, H=512): winner-mask agreement {agree}/{n_ok}")


def test_config_d_tango_128x128_1024_hypotheses():
    import spe_b200

    model = spe_b200.models.tango()
    fr = spe_b200.synth.make_frames(model, 64, 128, 128, seed=spe_b200.synth.BASE_SEED + 3)
    agree, n_ok = _stage_vs_oracle(model, fr, hypotheses=1024, min_agree=0.95)
    print(f"config D (J=11, 128x128, H=1024): winner-mask agreement {agree}/{n_ok}")


def test_full_batch_properties_config_b():
    """4096 x 11 x 64 x 64, 256 hypotheses: determinism, partition invariance (a frame's result
    does not depend on which batch/shard it is in), chunked host path == device path, unit
    quaternions, inliers are a subset of the visible landmarks."""
    import torch

    import spe_b200
    from spe_b200.pipeline import HeatmapToPose, shard_bounds

    model = spe_b200.models.tango()
    B = 4096
    hm, c, s = spe_b200.synth.device_heatmaps(model, B, 64, 64, seed=77, device="cuda")
    stage = HeatmapToPose(model, hypotheses=256)
    a = stage(hm, c, s)
    pose_a, mask_a, status_a = a.pose7.clone(), a.inlier_mask.clone(), a.status.clone()
    b = stage(hm, c, s)
    assert torch.equal(pose_a, b.pose7) and torch.equal(mask_a, b.inlier_mask) and torch.equal(status_a, b.status)
    for world in (2, 3):
        for r in range(world):
            lo, hi = shard_bounds(B, world, r)
            part = stage(hm[lo:hi].contiguous(), c[lo:hi].contiguous(), s[lo:hi].contiguous())
            assert torch.equal(part.pose7, pose_a[lo:hi]) and torch.equal(part.inlier_mask, mask_a[lo:hi])
    host = stage.run_host(hm.cpu().numpy(), c.cpu().numpy(), s.cpu().numpy(), chunk=600)  # ragged last chunk
    np.testing.assert_array_equal(host.pose7, pose_a.cpu().numpy())
    np.testing.assert_array_equal(host.status, status_a.cpu().numpy())
    ok = status_a == 0
    assert ok.float().mean() > 0.97
    q = pose_a[ok][:, :4].double().norm(dim=1)
    assert torch.allclose(q, torch.ones_like(q), atol=1e-6)
    kp_conf = a.kpts[..., 2]
    vis = (kp_conf > 1.94e-10).int()
    vis_bits = (vis << torch.arange(11, device="cuda", dtype=torch.int32)).sum(1).int()
    assert torch.all((mask_a & ~vis_bits) == 0)
    assert torch.all(pose_a[~ok] == 0)
    # translation sanity: the synthetic spacecraft sits 4-10 m in front of the camera
    tz = pose_a[ok][:, 6]
    assert (tz > 2).float().mean() > 0.99


def test_full_batch_config_c_decode_property():
    """16384 x 17 x 96 x 72 (7.7 GB of heatmaps): argmax equals torch.argmax on every map."""
    import torch

    import spe_b200

    B, J, H, W = 16384, 17, 96, 72
    g = torch.Generator(device="cuda").manual_seed(3)
    hm = torch.empty((B, J, H, W), device="cuda")
    for lo in range(0, B, 2048):
        hm[lo:lo + 2048] = torch.randn((2048, J, H, W), generator=g, device="cuda")
    p, m, idx = spe_b200.get_max_preds(hm, return_index=True)
    for lo in range(0, B, 2048):
        ref = hm[lo:lo + 2048].view(2048, J, -1).argmax(2)
        assert torch.equal(idx[lo:lo + 2048].long(), ref)
        assert torch.equal(m[lo:lo + 2048, :, 0], hm[lo:lo + 2048].view(2048, J, -1).amax(2))


def test_streamed_executor_matches_single_calls():
    """The software-pipelined executor (front of batch i+1 overlapping the tail of batch i on a
    side stream) returns exactly what back-to-back single calls return, slot reuse included."""
    import torch

    import spe_b200
    from spe_b200.pipeline import HeatmapToPose, StreamedHeatmapToPose

    model = spe_b200.models.tango()
    stage = HeatmapToPose(model, hypotheses=128)
    batches = [spe_b200.synth.device_heatmaps(model, 256, 64, 64, seed=100 + i, device="cuda") for i in range(5)]
    expect = []
    for hm, c, s in batches:
        o = stage(hm, c, s)
        expect.append((o.pose7.clone(), o.inlier_mask.clone(), o.status.clone(), o.kpts.clone()))
    pipe = StreamedHeatmapToPose(stage, 256, depth=2, want_rt=True)
    got = []
    for hm, c, s in batches:
        slot = pipe.submit(hm, c, s)
        pipe.wait(slot)  # results of a slot are only valid until it is reused
        o = slot["out"]
        got.append((o.pose7.clone(), o.inlier_mask.clone(), o.status.clone(), o.kpts.clone()))
    # and without synchronising in between (the executor must protect its own buffers)
    last = None
    for hm, c, s in batches:
        last = pipe.submit(hm, c, s)
    pipe.drain()
    torch.cuda.synchronize()
    # The background tail is the same kernel in another launch shape (8 warps per CTA instead of 1); poses
    # agree to ~1e-12 on R|t.  Masks, status and keypoints are identical.
    # Frames whose winner has exactly 5 inliers are excluded from the tight bound: EPnP on 5 points
    # has a 2-D null space and amplifies a 1e-16 perturbation to ~1e-4 (the same chaos that makes
    # per-hypothesis parity with cv2 statistical).
    def close(a, b, mask):
        five = torch.tensor([bin(int(v) & 0xFFFFFFFF).count("1") == 5 for v in mask.cpu()], device=a.device)
        d = (a - b).abs().amax(dim=1)
        return bool((d[~five] <= 2e-6).all()) and bool((d[five] <= 1e-2).all())

    for e, g in zip(expect, got):
        assert close(e[0], g[0], e[1])
        for a, b in zip(e[1:], g[1:]):
            assert torch.equal(a, b)
    assert close(last["out"].pose7, expect[-1][0], expect[-1][1])
    # the executor itself is deterministic
    again = pipe.submit(*batches[-1])
    pipe.drain()
    torch.cuda.synchronize()
    assert torch.equal(again["out"].pose7, last["out"].pose7)
    # caller-provided outputs: views of one [K*B, ...] buffer (what bench.py all_gathers once at the end of the job)
    from spe_b200.pipeline import StageOutput

    K = len(batches)
    big = StageOutput(torch.empty((K * 256, 7), device="cuda"), torch.empty((K * 256,), dtype=torch.int32, device="cuda"),
                      torch.empty((K * 256,), dtype=torch.int32, device="cuda"), None)
    for k, (hm, c, s) in enumerate(batches):
        sl = slice(k * 256, (k + 1) * 256)
        pipe.submit(hm, c, s, out=StageOutput(big.pose7[sl], big.inlier_mask[sl], big.status[sl], None))
    pipe.drain()
    torch.cuda.synchronize()
    for k, e in enumerate(expect):
        sl = slice(k * 256, (k + 1) * 256)
        assert torch.equal(big.inlier_mask[sl], e[1]) and torch.equal(big.status[sl], e[2]) and close(big.pose7[sl], e[0], e[1])
    # a caller running under its own stream: submit() follows torch's current stream
    own = torch.cuda.Stream()
    own.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(own):
        slot = pipe.submit(*batches[0])
        pipe.drain()
    own.synchronize()
    assert torch.equal(slot["out"].inlier_mask, expect[0][1])
