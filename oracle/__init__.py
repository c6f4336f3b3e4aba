"""CPU oracle for the heatmap->pose stage.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from this package; the product (spacecraft-pose-estimation_b200/) never does and
fails loudly when its CUDA library is missing.

Parity pin:
  * decode half — NumPy restatement of the reference's own functions
    (landmark_regression/lib/core/inference.py:18-79, lib/utils/transforms.py:49-110), pinned by
    golden vectors generated in the build container by importing the *reference itself*
    (tests/golden/make_golden.py -> tests/golden/decode_golden.npz) plus SURVEY App. E.1.
  * pose half — the arithmetic lives in an un-vendored third-party dependency, OpenCV calib3d
    (`opencv-python==3.4.11.41`, environment.yml:37; this image ships cv2 4.13.0, which is the
    executable oracle).  `pnp_ref.solve_pnp_ransac_cv2` calls cv2.solvePnPRansac with the
    reference's exact arguments (pose_estimation/export_predicted_poses_real.py:199-201);
    `pnp_ref.ransac_epnp_whitebox` restates OpenCV's RANSAC loop (fixed-seed RNG, 5-point minimal
    sets, adaptive iteration budget, final EPnP on the inliers) and is asserted bit-identical to
    the black box; `epnp_ref.epnp` restates EPnP itself in NumPy float64 and is asserted against
    cv2.solvePnP(EPNP).  The reference has no tests of its own for this path; golden PnP vectors
     were produced by cv2 4.13.0.
"""
