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


def get_final_preds_combined(config, heatmaps, center, scale, mode: str = "mean", flip_pairs=(), shift_heatmap: bool = False,
                             return_index: bool = False):
    """get_final_preds of a COMBINATION of heatmap tensors, computed while decoding — the combined
    tensor is never written.

    mode="mean": `heatmaps` is the list of K model outputs of the reference's ensemble validation
        (validate_cv, lib/core/function.py:525-536): decodes ((o0 + o1) + ...) / K.
    mode="flip": `heatmaps` = [output, output_flipped] where output_flipped is the network output
        on the mirrored input, exactly what the reference has before `flip_back`
        (lib/core/function.py:347-366): decodes (output + shift(flip_back(output_flipped))) * 0.5
        with `flip_pairs` = val_dataset.flip_pairs and shift_heatmap = config.TEST.SHIFT_HEATMAP.
    Same return conventions as get_final_preds.
    """
    import ctypes

    torch = _lib.require_cuda()
    L = _lib.lib()
    hms = list(heatmaps)
    if mode not in ("mean", "flip"):
        raise ValueError("mode must be 'mean' or 'flip'")
    if mode == "flip" and len(hms) != 2:
        raise ValueError("flip mode takes [output, output_flipped]")
    if not (1 <= len(hms) <= 8):
        raise ValueError("between 1 and 8 tensors")
    is_np = _check_heatmaps(hms[0], torch)
    dev = _device_of(hms[0], torch)
    dhm = [_to_device(h, torch, dev, "heatmaps") for h in hms]
    B, J, H, W = dhm[0].shape
    for h in dhm:
        if tuple(h.shape) != (B, J, H, W):
            raise ValueError("all tensors need the same shape")
    c = _to_device(center, torch, dev, "center").reshape(-1, 2)
    s = _to_device(scale, torch, dev, "scale").reshape(-1, 2)
    if c.shape[0] != B or s.shape[0] != B:
        raise ValueError("center and scale need one row per frame")
    perm = None
    if mode == "flip" and len(flip_pairs) > 0:
        p = np.arange(J, dtype=np.int32)
        for a, b in flip_pairs:
            p[a], p[b] = p[b], p[a]
        perm = torch.from_numpy(p).to(dev)
    kpts = torch.empty((B, J, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((B, J), dtype=torch.int32, device=dev) if return_index else None
    ptrs = (ctypes.c_void_p * len(dhm))(*[h.data_ptr() for h in dhm])
    with torch.cuda.device(dev):
        _lib.check(L.spe_decode_combined_kpts_f32(ptrs, len(dhm), 0 if mode == "mean" else 1, perm.data_ptr() if perm is not None else None,
                                                  int(bool(shift_heatmap)), B, J, H, W, c.data_ptr(), s.data_ptr(),
                                                  int(_post_process_flag(config)), kpts.data_ptr(), idx.data_ptr() if idx is not None else None,
                                                  torch.cuda.current_stream(dev).cuda_stream), "spe_decode_combined_kpts_f32")
    preds, maxvals = kpts[..., :2].contiguous(), kpts[..., 2:].contiguous()
    if is_np:
        out = (preds.cpu().numpy(), maxvals.cpu().numpy())
        return out + (idx.cpu().numpy(),) if return_index else out
    return (preds, maxvals, idx) if return_index else (preds, maxvals)
