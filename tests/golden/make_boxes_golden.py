"""Golden vectors for the box stage, produced by EXECUTING THE REFERENCE'S OWN SOURCE LINES here.

Run in the build container only (needs /root/reference):
    python tests/golden/make_boxes_golden.py
Writes tests/golden/boxes_golden.npz.

* `_xywh2cs` is taken from landmark_regression/lib/dataset/PEdataset.py (the `def _xywh2cs` block; the module itself
  cannot be imported here: json_tricks is absent and np.float is gone) and called with a stand-in `self` that carries
  the class's own `pixel_std` (parsed from the same file).
* the per-image box choice is the statement block of object_detection/export_object_detection_bounding_boxes.py between
  `output_box = None` and `bounding_box = [x, y, w, h]`, executed on seeded (boxes, scores) lists.
Nothing of the reference is copied into the repository: only inputs and outputs are stored.
"""
import os
import re
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def reference_xywh2cs():
    src = open(os.path.join(REF, "landmark_regression/lib/dataset/PEdataset.py")).read()
    m = re.search(r"^    def _xywh2cs\(self.*?^        return center, scale\n", src, flags=re.S | re.M)
    pixel_std = int(re.search(r"self\.pixel_std\s*=\s*(\d+)", src).group(1))
    ns = {"np": np}
    exec(textwrap.dedent(m.group(0)), ns)
    me = types.SimpleNamespace(pixel_std=pixel_std, aspect_ratio=1.0)
    return lambda x, y, w, h: ns["_xywh2cs"](me, x, y, w, h)


def reference_pick():
    src = open(os.path.join(REF, "object_detection/export_object_detection_bounding_boxes.py")).read()
    m = re.search(r"^        output_box = None\n.*?^        bounding_box = \[x, y, w, h\]\n", src, flags=re.S | re.M)
    code = compile(textwrap.dedent(m.group(0)), "export_object_detection_bounding_boxes.py[box choice]", "exec")

    def pick(boxes, scores, image_width, image_height):
        ns = {"np": np, "boxes": boxes, "scores": scores, "args": types.SimpleNamespace(image_width=image_width, image_height=image_height),
              "output_dir0": "", "output_dir1": "", "output_dir2": ""}
        exec(code, ns)
        return ns["bounding_box"], ns["output_score"]

    return pick


def main():
    rng = np.random.default_rng(20261017)
    xywh2cs, pick = reference_xywh2cs(), reference_pick()
    B, K = 96, 4
    boxes = np.zeros((B, K, 4), np.float32)
    scores = np.zeros((B, K), np.float32)
    counts = rng.integers(0, K + 1, B).astype(np.int32)
    counts[:8] = [0, 1, 2, 3, 4, 1, 2, 2]
    sizes = [(1920, 1200), (640, 480)]
    size_id = rng.integers(0, 2, B).astype(np.int32)
    for b in range(B):
        W, H = sizes[size_id[b]]
        x1 = rng.uniform(0, W * 0.7, K)
        y1 = rng.uniform(0, H * 0.7, K)
        boxes[b] = np.stack([x1, y1, x1 + rng.uniform(4, W * 0.3, K), y1 + rng.uniform(4, H * 0.3, K)], 1).astype(np.float32)
        scores[b] = rng.uniform(0.05, 1.0, K).astype(np.float32)
    scores[6, 1] = scores[6, 0]  # tie between two detections: the first wins
    scores[7, 1] = np.nan  # np.argmax: a NaN is the maximum
    scores[5, 0] = np.nan
    counts[20], scores[20, 0] = 2, np.nan  # NaN first: stays
    xywh = np.zeros((B, 4), np.float64)
    best = np.zeros(B, np.float64)
    center = np.zeros((B, 2), np.float32)
    scale = np.zeros((B, 2), np.float32)
    for b in range(B):
        W, H = sizes[size_id[b]]
        n = int(counts[b])
        bb, sc = pick(boxes[b, :n].copy(), scores[b, :n].copy(), W, H)
        xywh[b], best[b] = bb, sc
        c, s = xywh2cs(*np.array(bb).flatten()[:4])
        center[b], scale[b] = c, s
    # _xywh2cs alone on float64 COCO boxes, incl. the center[0] == -1 branch
    q = np.stack([rng.uniform(-50, 1900, 64), rng.uniform(-50, 1190, 64), rng.uniform(1, 900, 64), rng.uniform(1, 900, 64)], 1)
    q[0] = [-2.0, 10.0, 2.0, 30.0]  # center[0] == -1: scale is NOT multiplied by 1.5
    q[1] = [-1.0, 5.0, 0.0, 8.0]
    q[2] = [0.1, 0.2, 1e-3, 1e4]
    cs = [xywh2cs(*row) for row in q]
    np.savez_compressed(os.path.join(HERE, "boxes_golden.npz"), boxes=boxes, scores=scores, counts=counts, size_id=size_id,
                        sizes=np.array(sizes, np.int32), xywh=xywh, best_score=best, center=center, scale=scale,
                        q_xywh=q, q_center=np.array([c for c, _ in cs], np.float32), q_scale=np.array([s for _, s in cs], np.float32))
    print("wrote boxes_golden.npz")


if __name__ == "__main__":
    main()
