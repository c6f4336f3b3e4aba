"""Detection boxes -> (center, scale) on the device.

Mirrors the two host steps the reference runs between detectron2 and HRNet's decode:
  * pick_boxes   object_detection/export_object_detection_bounding_boxes.py:313-329 — one box per image: the best of
                 one or two detections, the whole image otherwise; leaves as the COCO [x, y, w, h] the script writes
  * xywh2cs      JointsDataset._xywh2cs, landmark_regression/lib/dataset/PEdataset.py:98-113 — the (center, scale)
                 pair `get_final_preds` takes
NumPy in -> NumPy out, torch CUDA in -> torch CUDA out; the arithmetic is in libspe_b200.so (csrc/boxes.cu).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def xywh2cs(xywh, device=None):
    """xywh [B,4] float64 -> (center [B,2] float32, scale [B,2] float32)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    as_numpy = isinstance(xywh, np.ndarray) or not hasattr(xywh, "is_cuda")
    t = torch.as_tensor(np.ascontiguousarray(xywh, np.float64)) if as_numpy else xywh
    t = t.to(device or ("cuda" if not t.is_cuda else t.device), torch.float64).contiguous().reshape(-1, 4)
    B = t.shape[0]
    center = torch.empty((B, 2), dtype=torch.float32, device=t.device)
    scale = torch.empty((B, 2), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(L.spe_boxes_to_center_scale_f64(t.data_ptr(), B, center.data_ptr(), scale.data_ptr(), torch.cuda.current_stream(t.device).cuda_stream),
                   "spe_boxes_to_center_scale_f64")
    return (center.cpu().numpy(), scale.cpu().numpy()) if as_numpy else (center, scale)


def pick_boxes(boxes, scores, counts, image_width, image_height, device=None):
    """boxes [B,K,4] float32 (x1,y1,x2,y2), scores [B,K] float32, counts [B] int (detections per image; None = K).
    Returns dict(xywh [B,4] float64, score [B] float32, index [B] int32, center [B,2], scale [B,2] float32)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    as_numpy = isinstance(boxes, np.ndarray)
    dev = torch.device(device or "cuda") if as_numpy else boxes.device
    bt = (torch.from_numpy(np.ascontiguousarray(boxes, np.float32)) if as_numpy else boxes).to(dev, torch.float32).contiguous()
    st = (torch.from_numpy(np.ascontiguousarray(scores, np.float32)) if as_numpy else scores).to(dev, torch.float32).contiguous()
    B, K = st.shape
    assert bt.shape == (B, K, 4)
    ct = None
    if counts is not None:
        ct = (torch.from_numpy(np.ascontiguousarray(counts, np.int32)) if isinstance(counts, np.ndarray) else counts).to(dev, torch.int32).contiguous()
    xywh = torch.empty((B, 4), dtype=torch.float64, device=dev)
    score = torch.empty((B,), dtype=torch.float32, device=dev)
    index = torch.empty((B,), dtype=torch.int32, device=dev)
    center = torch.empty((B, 2), dtype=torch.float32, device=dev)
    scale = torch.empty((B, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.spe_pick_boxes_f32(bt.data_ptr() if K else None, st.data_ptr() if K else None, ct.data_ptr() if ct is not None else None, B, K,
                                        float(image_width), float(image_height), xywh.data_ptr(), score.data_ptr(), index.data_ptr(),
                                        center.data_ptr(), scale.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "spe_pick_boxes_f32")
    out = dict(xywh=xywh, score=score, index=index, center=center, scale=scale)
    return {k: v.cpu().numpy() for k, v in out.items()} if as_numpy else out
