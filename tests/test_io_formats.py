"""pred.mat / opencv_poses.json writers: byte-level structure the reference's
own reader code expects (export_predicted_poses_real.py:172-173, :224-236)."""
import json

import numpy as np

from spe_b200 import io as spe_io


def test_pred_mat_round_trip_like_the_reference_reader(tmp_path):
    import scipy.io as scio

    rng = np.random.default_rng(0)
    kpts = rng.normal(size=(7, 11, 3)).astype(np.float32)
    p = spe_io.save_pred_mat(str(tmp_path / "results" / "pred"), kpts)
    assert p.endswith("pred.mat")
    # the reference: preds = scio.loadmat(args.pose_annotations); preds = np.array(preds['preds'])
    preds = np.array(scio.loadmat(p)["preds"])
    assert preds.dtype == np.float32 and preds.shape == (7, 11, 3)
    np.testing.assert_array_equal(preds, kpts)
    np.testing.assert_array_equal(spe_io.load_pred_mat(p), kpts)
    # and the reference's per-frame unpacking
    row = np.array(preds[3].flatten()).reshape((-1, 3))
    np.testing.assert_array_equal(row[:, :2].astype(np.float32), kpts[3, :, :2])


def test_opencv_poses_json_structure(tmp_path):
    rt = np.zeros((2, 12))
    rt[0, :9] = np.eye(3).ravel()
    rt[0, 9:] = [0.1, -0.2, 6.0]
    rt[1, :9] = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], float).ravel()
    rt[1, 9:] = [1.0, 2.0, 3.0]
    p = spe_io.save_opencv_poses_json(str(tmp_path / "opencv_poses.json"), ["img000001.jpg", "img000002.jpg"], rt, status=[0, 0])
    poses = json.load(open(p))
    assert [sorted(r.keys()) for r in poses] == [["T", "image_name", "rotation_matrix"]] * 2
    assert poses[0]["T"] == [[0.1], [-0.2], [6.0]]  # pred_T.tolist() of a (3,1) array
    assert poses[1]["rotation_matrix"] == [[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]
    assert open(p).read().startswith("[\n  {\n    \"image_name\"")  # json.dumps(indent=2)
    # a frame without a pose keeps the reference's keys and adds a status
    p2 = spe_io.save_opencv_poses_json(str(tmp_path / "b.json"), ["a.jpg"], np.zeros((1, 12)), status=[3])
    assert json.load(open(p2))[0]["status"] == 3
