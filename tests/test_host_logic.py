"""Host-side logic that needs no GPU: the frame sharding + final all_gather (world_size 2 over
gloo), the synthetic workload generator, the cv2-compatible rvec conversion."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    from spe_b200.pipeline import shard_bounds

    for n in (0, 1, 7, 64, 4096, 65537):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            sizes = [hi - lo for lo, hi in edges]
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
    from spe_b200.pipeline import all_gather_rows, shard_bounds

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, world, rank)
    full = torch.arange(n_total * 7, dtype=torch.float32).reshape(n_total, 7)
    got = all_gather_rows(full[lo:hi].clone(), n_total)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 9, 4097])
def test_all_gather_rows_world2_gloo(n_total):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + n_total) % 300
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_synth_is_seeded_and_well_formed():
    from spe_b200 import models, synth

    m = models.tango()
    a = synth.make_frames(m, 6, 64, 64, seed=5)
    b = synth.make_frames(m, 6, 64, 64, seed=5)
    c = synth.make_frames(m, 6, 64, 64, seed=6)
    np.testing.assert_array_equal(a.heatmaps, b.heatmaps)
    assert not np.array_equal(a.heatmaps, c.heatmaps)
    assert a.heatmaps.dtype == np.float32 and a.heatmaps.shape == (6, 11, 64, 64)
    assert a.center.dtype == np.float32 and a.scale.dtype == np.float32
    # every true landmark lands inside its heatmap
    clean = ~(a.outlier | a.masked)
    assert (a.peak_hm[clean] >= 0).all() and (a.peak_hm[clean][:, 0] <= 63).all() and (a.peak_hm[clean][:, 1] <= 63).all()
    assert (a.heatmaps[a.masked].max(axis=(-1, -2)) <= 0).all() if a.masked.any() else True
    # the projection formula is cv2.projectPoints
    import cv2

    p_cv, _ = cv2.projectPoints(m.landmarks, a.rvec[0], a.tvec[0], m.K, m.dist)
    np.testing.assert_allclose(a.image_points[0], p_cv.reshape(-1, 2), atol=1e-8)
    h = models.hubble_synthetic(17)
    assert h.num_landmarks == 17 and models.hubble_synthetic(24).num_landmarks == 24


def test_matrix_to_rvec_matches_cv2():
    import cv2

    from spe_b200.pnp import matrix_to_rvec

    rng = np.random.default_rng(0)
    for _ in range(100):
        rv = rng.normal(size=3)
        rv *= rng.uniform(0, np.pi - 1e-3) / np.linalg.norm(rv)
        R = cv2.Rodrigues(rv)[0]
        np.testing.assert_allclose(matrix_to_rvec(R), cv2.Rodrigues(R)[0].ravel(), atol=1e-9)
    np.testing.assert_allclose(matrix_to_rvec(np.eye(3)), 0, atol=1e-15)


def test_bench_reference_arm_runs_on_cpu():
    import json
    import subprocess

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "frames/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0
