"""Spacecraft landmark models and camera calibrations used by the heatmap->pose stage.

These are *data*, not code: the Tango numbers are the reference's own fixtures
(object_detection/speed_plus_utils/landmarks.csv — 11 landmarks in metres — and
object_detection/speed_plus_utils/calibration.json — SPEED+ camera, 1920x1200), restated here
because /root/reference does not exist on the GPU box.  The Hubble event-camera model is not
shipped by the reference (pose_estimation/projection_utils/landmarks_hubble.csv is referenced by
train_pipeline_hubble_dvx.sh:39 but absent), so `hubble_synthetic` follows SURVEY.md §8(d): seeded
uniform landmarks and an undistorted 640x480 pinhole (the pipeline undistorts event frames first,
v2e/convert_aedats.py:56-60).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class CameraModel:
    """Landmarks [J,3] float64 (metres), K [3,3] float64, dist [5] float64 = (k1,k2,p1,p2,k3)."""

    name: str
    landmarks: np.ndarray
    K: np.ndarray
    dist: np.ndarray
    image_size: tuple  # (width, height) in pixels

    @property
    def num_landmarks(self) -> int:
        return int(self.landmarks.shape[0])


_TANGO_LANDMARKS = np.array(
    [
        [0.36940446496009827, -0.3845726549625397, 0.16007566452026367],
        [0.36786314845085144, 0.3836139440536499, 0.16053038835525513],
        [-0.36881211400032043, 0.38277047872543335, 0.16048267483711243],
        [-0.36801040172576904, -0.3831963539123535, 0.16058564186096191],
        [0.36815810203552246, -0.26237574219703674, -0.16152474284172058],
        [0.36859363317489624, 0.30254653096199036, -0.15993139147758484],
        [-0.36717548966407776, 0.30379965901374817, -0.1599225401878357],
        [-0.3663908839225769, -0.2586885094642639, -0.1586388796567917],
        [0.30565211176872253, -0.5800656676292419, 0.08969831466674805],
        [0.5425941348075867, 0.48880907893180847, 0.09245043992996216],
        [-0.5449637770652771, 0.48740869760513306, 0.09220433235168457],
    ],
    dtype=np.float64,
)

_SPEEDPLUS_K = np.array(
    [[2988.5795163815555, 0.0, 960.0], [0.0, 2988.3401159176124, 600.0], [0.0, 0.0, 1.0]],
    dtype=np.float64,
)

_SPEEDPLUS_DIST = np.array(
    [
        -0.22383016606510672,
        0.51409797089106379,
        -0.00066499611998340662,
        -0.00021404771667484594,
        -0.13124227429077406,
    ],
    dtype=np.float64,
)


def tango() -> CameraModel:
    """SPEED+ Tango: 11 landmarks, SPEED+ camera (configs A, B, D of BASELINE.json)."""
    return CameraModel("tango", _TANGO_LANDMARKS.copy(), _SPEEDPLUS_K.copy(), _SPEEDPLUS_DIST.copy(), (1920, 1200))


def hubble_synthetic(num_landmarks: int = 17) -> CameraModel:
    """Synthetic stand-in for the (unshipped) Hubble/DVX model, config C of BASELINE.json.

    J = 17 per experiments/events/events-config.yaml:28 (24 with the shipped overrides,
    train_pipeline_hubble_dvx.sh:53).  Landmarks ~ U([-1,1]^3)*(0.6,0.6,0.3) m seeded with J.
    """
    rng = np.random.default_rng(num_landmarks)
    lm = rng.uniform(-1.0, 1.0, size=(num_landmarks, 3)) * np.array([0.6, 0.6, 0.3])
    K = np.array([[600.0, 0.0, 320.0], [0.0, 600.0, 240.0], [0.0, 0.0, 1.0]])
    return CameraModel(f"hubble_synth{num_landmarks}", lm, K, np.zeros(5), (640, 480))
