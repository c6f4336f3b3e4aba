"""Oracle (test infrastructure): the two host steps between detectron2's boxes and the decode, restated.

* `pick_box` — object_detection/export_object_detection_bounding_boxes.py:313-329: per image, exactly one or two
  detections keep `boxes[scores.argmax()]`; any other count falls back to `[[0, 0, image_width, image_height]]` with
  score 0.  `.tolist()` turns the float32 corners into Python floats BEFORE the width/height subtraction, so the COCO
  `bbox = [x, y, w, h]` is float64.
* `xywh2cs` — JointsDataset._xywh2cs, landmark_regression/lib/dataset/PEdataset.py:98-113 with pixel_std = 200 (:41):
  float32 `center`, float32 `scale` multiplied by 1.5 unless `center[0] == -1`.

Pinned by tests/golden/boxes_golden.npz, produced by executing the reference's own source lines
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

PIXEL_STD = 200


def pick_box(boxes: np.ndarray, scores: np.ndarray, image_width, image_height):
    """boxes [k,4] float32 xyxy, scores [k] float32 -> ([x, y, w, h] Python floats, score)."""
    if len(boxes) not in (1, 2):
        boxes = np.array([[0, 0, image_width, image_height]])
        scores = np.array([0])
    output_box = boxes[scores.argmax()].tolist()
    output_score = scores.max()
    x = output_box[0]
    y = output_box[1]
    w = output_box[2] - output_box[0]
    h = output_box[3] - output_box[1]
    return [x, y, w, h], output_score


def xywh2cs(x, y, w, h):
    center = np.zeros((2), dtype=np.float32)
    center[0] = x + w * 0.5
    center[1] = y + h * 0.5
    scale = np.array([w * 1.0 / PIXEL_STD, h * 1.0 / PIXEL_STD], dtype=np.float32)
    if center[0] != -1:
        scale = scale * 1.5
    return center, scale
