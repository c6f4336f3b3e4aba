"""Drop-in replacements for the reference's heatmap decoding, running on the B200.

Same names, argument meaning and return conventions as
    get_max_preds(batch_heatmaps)                          landmark_regression/lib/core/inference.py:18-46
    get_final_preds(config, batch_heatmaps, center, scale) landmark_regression/lib/core/inference.py:49-79
with one extension: besides numpy.ndarray the functions take torch CUDA tensors and then return
torch CUDA tensors, so the caller (`validate`, lib/core/function.py:389) can drop its
`output.clone().cpu().numpy()` round trip.  NumPy in -> NumPy out (host buffers are copied to
the device and the results copied back), exactly the reference's types and shapes:
preds float32 [B,J,2], maxvals float32 [B,J,1].

All arithmetic happens in libspe_b200.so (csrc/decode.cu) through the C ABI of
include/spe_b200.h; there is no CPU path here.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _post_process_flag(config) -> bool:
    """The reference reads exactly one thing from its config: config.TEST.POST_PROCESS
    (inference.py:56).  A plain bool is accepted too."""
    if isinstance(config, (bool, np.bool_)):
        return bool(config)
    return bool(config.TEST.POST_PROCESS)


def _to_device(x, torch, device, what):
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return t.to(device, non_blocking=False)
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.float32:
            x = x.float()
        return x.to(device).contiguous()
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=device)


def _check_heatmaps(batch_heatmaps, torch):
    is_np = isinstance(batch_heatmaps, np.ndarray)
    assert is_np or isinstance(batch_heatmaps, torch.Tensor), "batch_heatmaps should be numpy.ndarray"
    assert batch_heatmaps.ndim == 4, "batch_images should be 4-ndim"
    return is_np


def _device_of(batch_heatmaps, torch):
    if isinstance(batch_heatmaps, torch.Tensor) and batch_heatmaps.is_cuda:
        return batch_heatmaps.device
    return torch.device("cuda", torch.cuda.current_device())


def decode_device(hm, center=None, scale=None, post_process=True, want_index=False, kpts_layout=False):
    """Device-resident core used by every public entry point.

    hm [B,J,H,W] float32 CUDA contiguous; center/scale [B,2] float32 CUDA or None (argmax only).
    Returns (preds [B,J,2], maxvals [B,J,1], index [B,J] int32 | None) or, with kpts_layout,
    (kpts [B,J,3], index | None).  Runs on torch's current stream; does not synchronise.
    """
    torch = _lib.require_cuda()
    L = _lib.lib()
    B, J, H, W = hm.shape
    assert hm.is_cuda and hm.dtype == torch.float32 and hm.is_contiguous()
    stream = torch.cuda.current_stream(hm.device).cuda_stream
    idx = torch.empty((B, J), dtype=torch.int32, device=hm.device) if want_index else None
    idx_ptr = idx.data_ptr() if idx is not None else None
    with torch.cuda.device(hm.device):
        if kpts_layout:
            kpts = torch.empty((B, J, 3), dtype=torch.float32, device=hm.device)
            _lib.check(L.spe_decode_kpts_f32(hm.data_ptr(), B, J, H, W, center.data_ptr(), scale.data_ptr(), int(post_process),
                                             kpts.data_ptr(), idx_ptr, stream), "spe_decode_kpts_f32")
            return kpts, idx
        preds = torch.empty((B, J, 2), dtype=torch.float32, device=hm.device)
        maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=hm.device)
        if center is None:
            _lib.check(L.spe_max_preds_f32(hm.data_ptr(), B, J, H, W, preds.data_ptr(), maxvals.data_ptr(), idx_ptr, stream),
                       "spe_max_preds_f32")
        else:
            _lib.check(L.spe_decode_f32(hm.data_ptr(), B, J, H, W, center.data_ptr(), scale.data_ptr(), int(post_process),
                                        preds.data_ptr(), maxvals.data_ptr(), idx_ptr, stream), "spe_decode_f32")
    return preds, maxvals, idx


def get_max_preds(batch_heatmaps, return_index: bool = False):
    """get predictions from score maps — heatmaps: [batch_size, num_joints, height, width]."""
    torch = _lib.require_cuda()
    is_np = _check_heatmaps(batch_heatmaps, torch)
    dev = _device_of(batch_heatmaps, torch)
    hm = _to_device(batch_heatmaps, torch, dev, "batch_heatmaps")
    preds, maxvals, idx = decode_device(hm, want_index=return_index)
    if is_np:
        out = (preds.cpu().numpy(), maxvals.cpu().numpy())
        return out + (idx.cpu().numpy(),) if return_index else out
    return (preds, maxvals, idx) if return_index else (preds, maxvals)


def get_final_preds(config, batch_heatmaps, center, scale, return_index: bool = False):
    """Heatmaps -> landmark coordinates in image pixels (argmax, quarter-pixel refinement when
    config.TEST.POST_PROCESS, inverse affine of each frame's detection box)."""
    torch = _lib.require_cuda()
    is_np = _check_heatmaps(batch_heatmaps, torch)
    dev = _device_of(batch_heatmaps, torch)
    hm = _to_device(batch_heatmaps, torch, dev, "batch_heatmaps")
    c = _to_device(center, torch, dev, "center").reshape(-1, 2)
    s = _to_device(scale, torch, dev, "scale").reshape(-1, 2)
    if c.shape[0] != hm.shape[0] or s.shape[0] != hm.shape[0]:
        raise ValueError("center and scale need one row per frame")
    preds, maxvals, idx = decode_device(hm, c, s, _post_process_flag(config), want_index=return_index)
    if is_np:
        out = (preds.cpu().numpy(), maxvals.cpu().numpy())
        return out + (idx.cpu().numpy(),) if return_index else out
    return (preds, maxvals, idx) if return_index else (preds, maxvals)
