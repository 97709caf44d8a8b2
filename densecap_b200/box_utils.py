"""Mirror of densecap/box_utils.lua (the functions on the test-time hot path), backed by the CUDA
kernels of libdensecap_b200.so.  Tensors are torch CUDA fp32 tensors; indices are 1-based int64,
exactly as the reference returns them."""
from __future__ import annotations

import torch

from . import _lib
from .runtime import default_context, ptr, stream_ptr, _check_cuda_f32


def xcycwh_to_x1y1x2y2(boxes):
    """box_utils.xcycwh_to_x1y1x2y2 (densecap/box_utils.lua:270-298)."""
    b = _check_cuda_f32(boxes)
    out = torch.empty_like(b)
    _lib.check(_lib.load().dc_xcycwh_to_x1y1x2y2(ptr(b), ptr(out), b.numel() // 4, stream_ptr()))
    return out


def x1y1x2y2_to_xcycwh(boxes):
    """box_utils.x1y1x2y2_to_xcycwh (densecap/box_utils.lua:382-410)."""
    b = _check_cuda_f32(boxes)
    out = torch.empty_like(b)
    _lib.check(_lib.load().dc_x1y1x2y2_to_xcycwh(ptr(b), ptr(out), b.numel() // 4, stream_ptr()))
    return out


def clip_boxes(boxes, bounds, format="xcycwh"):
    """box_utils.clip_boxes (densecap/box_utils.lua:486-523).  bounds: dict x_min,x_max,y_min,y_max.
    Returns (boxes_clipped, valid uint8)."""
    if format != "xcycwh":
        raise ValueError('Unrecognized box format "%s"' % format)   # only the hot-path format is built
    b = _check_cuda_f32(boxes)
    n = b.numel() // 4
    out = torch.empty_like(b)
    valid = torch.empty((n,), dtype=torch.uint8, device=b.device)
    _lib.check(_lib.load().dc_clip_boxes(ptr(b), ptr(out), ptr(valid), n, float(bounds["x_min"]), float(bounds["y_min"]),
                                         float(bounds["x_max"]), float(bounds["y_max"]), stream_ptr()))
    return out, valid


def nms(boxes, overlap, max_boxes=None, ctx=None):
    """box_utils.nms (densecap/box_utils.lua:154-256).  boxes: (N,5) x1,y1,x2,y2,score.
    Returns 1-based int64 indices in decreasing-score order (K <= max_boxes)."""
    b = _check_cuda_f32(boxes)
    n = b.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=b.device)
    assert b.dim() == 2 and b.shape[1] == 5
    ctx = ctx or default_context(b.device.index or 0, min_boxes=n)
    cap = n if not max_boxes or max_boxes <= 0 else min(n, max_boxes)
    idx = torch.empty((max(cap, 1),), dtype=torch.int64, device=b.device)
    count = torch.zeros((1,), dtype=torch.int32, device=b.device)
    _lib.check(_lib.load().dc_nms(ctx.handle, ptr(b), n, float(overlap), int(max_boxes or 0), ptr(idx), ptr(count), stream_ptr()),
               ctx.handle)
    return idx[: int(count.item())]
