"""Drop-in replacement for the reference's training / validation metric, running on the B200.

    accuracy(output, target, hm_type='gaussian', thr=0.5)      landmark_regression/lib/core/evaluate.py:42-80
as called from lib/core/function.py:61-62 (train) and :395-396 (validate) with the network output and the target
heatmaps.  Same return tuple: (acc [J+1] float64 NumPy with acc[0] the mean, avg_acc, cnt, pred [B,J,2] float32).

The reference copies BOTH heatmap tensors to the host for this number (`output.detach().cpu().numpy()`,
`target.detach().cpu().numpy()`).  Here both argmaxes run on the device (spe_max_preds_f32, the decode kernel without the
affine) and so does the per-joint counting (spe_pck_counts_f32, csrc/evaluate.cu); 8 bytes per joint come back.
torch CUDA tensors in -> `pred` is a torch CUDA tensor; NumPy in -> NumPy out, exactly the reference's types.

Like the reference, `thr` is accepted and NOT used: accuracy() calls dist_acc(dists[idx[i]]) without it (evaluate.py:71),
so every joint is tested against dist_acc's own default 0.5.  There is no CPU path here.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .inference import _check_heatmaps, _device_of, _to_device, decode_device

_DIST_ACC_THR = 0.5  # dist_acc's default (evaluate.py:32), the only threshold the reference ever applies


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    torch = _lib.require_cuda()
    if hm_type != "gaussian":
        raise ValueError("hm_type must be 'gaussian' (the reference defines `pred` for that type only)")
    is_np = _check_heatmaps(output, torch)
    _check_heatmaps(target, torch)
    assert tuple(output.shape) == tuple(target.shape), "output and target heatmaps differ in shape"
    dev = _device_of(output, torch)
    out_d, tgt_d = _to_device(output, torch, dev, "output"), _to_device(target, torch, dev, "target")
    B, J, H, W = out_d.shape
    with torch.cuda.device(dev):
        pred, _, _ = decode_device(out_d)
        tgt, _, _ = decode_device(tgt_d)
        counts = torch.empty((J, 2), dtype=torch.int32, device=dev)
        # norm = ones * [h, w] / 10 divides (x, y): x by H / 10 and y by W / 10, the reference's order (evaluate.py:55-58)
        _lib.check(_lib.lib().spe_pck_counts_f32(pred.data_ptr(), tgt.data_ptr(), B, J, float(np.float64(H) / 10), float(np.float64(W) / 10),
                                                 _DIST_ACC_THR, counts.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "spe_pck_counts_f32")
        c = counts.cpu().numpy()
    acc = np.zeros((J + 1))
    avg_acc, cnt = 0, 0
    for i in range(J):
        valid, below = int(c[i, 0]), int(c[i, 1])
        acc[i + 1] = below * 1.0 / valid if valid > 0 else -1  # dist_acc
        if acc[i + 1] >= 0:
            avg_acc = avg_acc + acc[i + 1]
            cnt += 1
    avg_acc = avg_acc / cnt if cnt != 0 else 0
    if cnt != 0:
        acc[0] = avg_acc
    return acc, avg_acc, cnt, (pred.cpu().numpy() if is_np else pred)
