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
    assert lib.spe_abi_version() == 2
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
    assert lib.spe_pnp_model_create(lm, 11, K, None, 16385, ctypes.byref(handle)) == -1  # above SPE_MAX_HYPOTHESES
    assert lib.spe_pnp_model_create(lm, 11, K, None, 10000, ctypes.byref(handle)) == -4  # K with skew / a bad last row: unsupported
    assert b"unsupported" in lib.spe_status_string(-4) and b"memory" in lib.spe_status_string(-5)
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
    assert lib.spe_ransac_replay_f64(None, 4, 64, 15.0, 0.99, None, 0, None) == -1
    assert lib.spe_ransac_read_budget(None, None, 4, 64, None, None) == -1
    assert lib.spe_pnp_model_num_landmarks(None) == -1
    ptrs = (ctypes.c_void_p * 2)(None, None)
    assert lib.spe_decode_combined_kpts_f32(ptrs, 0, 0, None, 0, 4, 11, 64, 64, None, None, 1, None, None, None) == -1  # K < 1
    # accuracy(): null coordinates, no joints, a non-positive divisor, coordinates that are not float2-aligned
    assert lib.spe_pck_counts_f32(None, None, 4, 11, 6.4, 6.4, 0.5, ctypes.c_void_p(16), None) == -1
    assert lib.spe_pck_counts_f32(ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 0, 6.4, 6.4, 0.5, ctypes.c_void_p(16), None) == -1
    assert lib.spe_pck_counts_f32(ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 11, 0.0, 6.4, 0.5, ctypes.c_void_p(16), None) == -1
    assert lib.spe_pck_counts_f32(ctypes.c_void_p(20), ctypes.c_void_p(16), 4, 11, 6.4, 6.4, 0.5, ctypes.c_void_p(16), None) == -1
    assert lib.spe_pck_counts_f32(ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 11, 6.4, 6.4, 0.5, None, None) == -1
    assert lib.spe_decode_combined_kpts_f32(ptrs, 9, 0, None, 0, 4, 11, 64, 64, None, None, 1, None, None, None) == -1  # K > 8
    assert lib.spe_decode_combined_kpts_f32(ptrs, 3, 1, None, 0, 4, 11, 64, 64, None, None, 1, None, None, None) == -1  # flip needs K == 2
    assert lib.spe_decode_combined_kpts_f32(ptrs, 2, 7, None, 0, 4, 11, 64, 64, None, None, 1, None, None, None) == -1  # unknown mode
    assert lib.spe_decode_combined_kpts_f32(ptrs, 2, 0, None, 0, 0, 11, 64, 64, None, None, 1, None, None, None) == 0  # empty batch


def test_control_point_table_entries_on_the_host(lib):
    """spe_pnp_control_entry (host only): the per-subset control-point data the hypothesis kernel looks up.
    The entry must reproduce the geometry of the five points — alphas are coordinates in an orthonormal PCA
    frame scaled by k_i = sqrt(lambda_i / 5) (OpenCV epnp.cpp choose_control_points) — and the rank must walk
    the table in the order model creation fills it."""
    import ctypes
    import itertools

    rng = np.random.default_rng(5)
    J = 9
    lm = np.ascontiguousarray(rng.uniform(-1, 1, (J, 3)) * np.array([0.6, 0.5, 0.2]))
    lm32 = lm.astype(np.float32).astype(np.float64)
    dp = lm.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    ranks = []
    for ids in itertools.combinations(range(J), 5):
        idv = np.array(ids, np.int32)
        entry = np.zeros(20, np.float32)
        rank = ctypes.c_int64(-1)
        rc = lib.spe_pnp_control_entry(dp, J, idv.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                       entry.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.byref(rank))
        assert rc == 0
        ranks.append(rank.value)
        al = entry[:15].reshape(5, 3).astype(np.float64)
        ksq = entry[15:18].astype(np.float64)
        assert np.all(entry[18:] == 0)
        P = lm32[list(ids)]
        # principal components: zero mean, orthogonal, variance 5 (unit-variance coordinates times k_i)
        np.testing.assert_allclose(al.sum(0), 0, atol=2e-5)
        np.testing.assert_allclose(al.T @ al, 5 * np.eye(3), atol=2e-4)
        # k_i^2 = eigenvalues of the covariance / 5
        lam = np.linalg.eigvalsh((P - P.mean(0)).T @ (P - P.mean(0))) / 5
        np.testing.assert_allclose(np.sort(ksq), lam, rtol=2e-5, atol=1e-9)
        # the alphas and k reproduce every pairwise distance of the five points
        d_true = ((P[:, None] - P[None]) ** 2).sum(-1)
        d_tab = (((al[:, None] - al[None]) ** 2) * ksq).sum(-1)
        np.testing.assert_allclose(d_tab, d_true, rtol=2e-4, atol=1e-7)
    # combinations in colexicographic order = 0 .. C(J,5)-1, each exactly once
    assert sorted(ranks) == list(range(126))
    colex = sorted(itertools.combinations(range(J), 5), key=lambda t: t[::-1])
    assert [ranks[list(itertools.combinations(range(J), 5)).index(t)] for t in colex] == list(range(126))
    # argument errors: unsorted / repeated / out-of-range ids
    entry = np.zeros(20, np.float32)
    for bad in ([0, 2, 1, 3, 4], [0, 1, 1, 3, 4], [0, 1, 2, 3, 9]):
        idv = np.array(bad, np.int32)
        assert lib.spe_pnp_control_entry(dp, J, idv.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                         entry.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), None) == -1


def test_minimal_sets_and_duplicate_free_lists_on_the_host(lib):
    """spe_pnp_minimal_sets_host (host only): what spe_pnp_model_create uploads for one point count.  The draws must be
    OpenCV's (oracle.ocv_rng, pinned by SURVEY App. E.2), `slot` must map every draw to the first draw of the same
    5-subset (as a set), `uniq` must list those first draws in ascending order — so that the distinct sets among the
    first H draws are a prefix of the list for every H — and the duplicate fractions must be the ones DESIGN.md quotes."""
    from oracle import ocv_rng

    def host_tables(n, H):
        sets, slot, uniq = np.zeros((H, 5), np.int32), np.zeros(H, np.int32), np.zeros(H, np.int32)
        nu = ctypes.c_int32()
        p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        assert lib.spe_pnp_minimal_sets_host(n, H, p(sets), p(slot), p(uniq), ctypes.byref(nu)) == 0
        return sets, slot, uniq, nu.value

    expected_unique = {(11, 256): 191, (11, 1024): 414, (11, 2048): 452}
    for n, H in ((6, 64), (10, 256), (11, 256), (11, 1024), (11, 2048), (17, 512), (24, 10000), (32, 4096)):
        sets, slot, uniq, nu = host_tables(n, H)
        np.testing.assert_array_equal(sets, ocv_rng.minimal_sets(n, H))
        seen = {}
        for h in range(H):
            key = tuple(sorted(sets[h]))
            if key not in seen:
                seen[key] = len(seen)
                assert uniq[seen[key]] == h
            assert slot[h] == seen[key]
        assert nu == len(seen) and np.all(uniq[nu:] == -1) and np.all(np.diff(uniq[:nu]) > 0)
        for cut in (1, 32, H // 2, H):  # prefix property: distinct sets among the first `cut` draws
            assert len({tuple(sorted(s)) for s in sets[:cut]}) == int(np.searchsorted(uniq[:nu], cut))
        if (n, H) in expected_unique:
            assert nu == expected_unique[(n, H)]
    assert lib.spe_pnp_minimal_sets_host(5, 64, None, None, None, None) == -1
    assert lib.spe_pnp_minimal_sets_host(11, 0, None, None, None, None) == -1
