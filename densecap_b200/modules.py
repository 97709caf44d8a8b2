"""Mirrors of the leaf nn.Modules under densecap/modules/ that sit on the test-time hot path.
Each class keeps the reference's constructor arguments, `forward` protocol and error behaviour, and
calls one entry point of libdensecap_b200.so (no Python arithmetic, no CPU fallback)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .runtime import default_context, ptr, stream_ptr, _check_cuda_f32


class MakeAnchors:
    """nn.MakeAnchors (densecap/modules/MakeAnchors.lua:31-67)."""

    def __init__(self, x0, y0, sx, sy, anchors):
        self.x0, self.y0, self.sx, self.sy = float(x0), float(y0), float(sx), float(sy)
        self.anchors = np.ascontiguousarray(np.asarray(anchors, np.float32))        # (2,k)
        assert self.anchors.ndim == 2 and self.anchors.shape[0] == 2

    def forward(self, input):
        N, H, W = input.shape[0], input.shape[2], input.shape[3]
        k = self.anchors.shape[1]
        out = torch.empty((N, 4 * k, H, W), dtype=torch.float32, device=input.device)
        _lib.check(_lib.load().dc_make_anchors(ptr(out), N, H, W, k, self.x0, self.y0, self.sx, self.sy,
                                               self.anchors.ctypes.data_as(C.c_void_p), stream_ptr()))
        return out


class ReshapeBoxFeatures:
    """nn.ReshapeBoxFeatures (densecap/modules/ReshapeBoxFeatures.lua:8-33)."""

    def __init__(self, k):
        self.k = int(k)

    def forward(self, input):
        x = _check_cuda_f32(input)
        N, kD, H, W = x.shape
        D = kD // self.k
        out = torch.empty((N, self.k * H * W, D), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().dc_reshape_box_features(ptr(x), ptr(out), N, self.k, D, H, W, stream_ptr()))
        return out


class ApplyBoxTransform:
    """nn.ApplyBoxTransform (densecap/modules/ApplyBoxTransform.lua:63-90).  forward({boxes, trans})."""

    def forward(self, input):
        boxes, trans = input
        assert boxes.shape[-1] == 4, "Last dim of boxes must be 4"
        assert trans.shape[-1] == 4, "Last dim of trans must be 4"
        b = _check_cuda_f32(boxes); t = _check_cuda_f32(trans)
        out = torch.empty_like(b)
        _lib.check(_lib.load().dc_apply_box_transform(ptr(b), ptr(t), ptr(out), b.numel() // 4, stream_ptr()))
        return out


class InvertBoxTransform:
    """nn.InvertBoxTransform (densecap/modules/InvertBoxTransform.lua:36-60).  forward({anchors, targets})."""

    def forward(self, input):
        a, t = (_check_cuda_f32(x) for x in input)
        out = torch.empty_like(t)
        _lib.check(_lib.load().dc_invert_box_transform(ptr(a), ptr(t), ptr(out), t.numel() // 4, stream_ptr()))
        return out


class BoxToAffine:
    """nn.BoxToAffine (densecap/modules/BoxToAffine.lua:55-93); setSize(H, W) before each forward."""

    def __init__(self):
        self.H = None
        self.W = None

    def setSize(self, H, W):
        self.H, self.W = int(H), int(W)
        return self

    def forward(self, input):
        assert input.dim() == 2, "Expected 2D input"
        assert input.shape[1] == 4, "Expected input of shape B x 4"
        assert self.H and self.W, "Need to call setSize before calling forward"
        b = _check_cuda_f32(input)
        out = torch.empty((b.shape[0], 2, 3), dtype=torch.float32, device=b.device)
        _lib.check(_lib.load().dc_box_to_affine(ptr(b), ptr(out), b.shape[0], self.H, self.W, stream_ptr()))
        return out


class AffineGridGeneratorBHWD:
    """stnbhwd nn.AffineGridGeneratorBHWD(height, width) (call site densecap/modules/BilinearRoiPooling.lua:52):
    forward(theta (B,2,3)) -> grids (B,height,width,2), last dimension (y, x)."""

    def __init__(self, height, width):
        self.height, self.width = int(height), int(width)

    def forward(self, theta):
        t = _check_cuda_f32(theta)
        assert t.dim() == 3 and t.shape[1] == 2 and t.shape[2] == 3, "Expected input of shape B x 2 x 3"
        out = torch.empty((t.shape[0], self.height, self.width, 2), dtype=torch.float32, device=t.device)
        _lib.check(_lib.load().dc_affine_grid_generator(ptr(t), t.shape[0], self.height, self.width, ptr(out), stream_ptr()))
        return out


class BatchBilinearSamplerBHWD:
    """nn.BatchBilinearSamplerBHWD (densecap/modules/BatchBilinearSamplerBHWD.lua:104-122): forward({feats (H,W,C), grids (B,HH,WW,2)})
    -> (B,HH,WW,C); one feature image shared by every grid."""

    def __init__(self, ctx=None):
        self.ctx = ctx

    def forward(self, input):
        feats, grids = input
        f = _check_cuda_f32(feats); g = _check_cuda_f32(grids)
        assert g.dim() == 4 and f.dim() == 3
        H, W, C_ = f.shape
        B, HH, WW, _ = g.shape
        ctx = self.ctx or default_context(f.device.index or 0)
        out = torch.empty((B, HH, WW, C_), dtype=torch.float32, device=f.device)
        _lib.check(_lib.load().dc_batch_bilinear_sampler(ctx.handle, ptr(f), H, W, C_, ptr(g), B, HH, WW, ptr(out), stream_ptr()), ctx.handle)
        return out


class BoxIoU:
    """nn.BoxIoU (densecap/modules/BoxIoU.lua:40-73).  forward({box1 (N,B1,4), box2 (N,B2,4)}) with N == 1."""

    def forward(self, input):
        b1, b2 = (_check_cuda_f32(x) for x in input)
        assert b1.shape[0] == 1 and b2.shape[0] == 1, "Only minibatches with N = 1 are supported"
        B1, B2 = b1.shape[1], b2.shape[1]
        out = torch.empty((1, B1, B2), dtype=torch.float32, device=b1.device)
        _lib.check(_lib.load().dc_box_iou(ptr(b1), ptr(b2), ptr(out), B1, B2, stream_ptr()))
        return out


class BilinearRoiPooling:
    """nn.BilinearRoiPooling (densecap/modules/BilinearRoiPooling.lua:35-91).
    forward({feats (C,H,W), boxes (B,4) xcycwh in image coordinates}) -> (B,C,height,width)."""

    def __init__(self, height, width, ctx=None):
        self.height, self.width = int(height), int(width)
        self.image_height = None
        self.image_width = None
        self.called_forward = False
        self.ctx = ctx

    def setImageSize(self, image_height, image_width):
        self.image_height, self.image_width = int(image_height), int(image_width)
        self.called_forward = False
        return self

    def forward(self, input, layout="BCHW"):
        assert self.image_height and self.image_width and not self.called_forward, \
            "Must call setImageSize before each forward pass"
        feats, boxes = input
        f = _check_cuda_f32(feats); b = _check_cuda_f32(boxes)
        C_, H, W = f.shape
        B = b.shape[0]
        ctx = self.ctx or default_context(f.device.index or 0)
        shape = (B, C_, self.height, self.width) if layout == "BCHW" else (B, self.height, self.width, C_)
        out = torch.empty(shape, dtype=torch.float32, device=f.device)
        _lib.check(_lib.load().dc_bilinear_roi_pool(ctx.handle, ptr(f), C_, H, W, ptr(b), B, self.image_height, self.image_width,
                                                    self.height, self.width, ptr(out), 0 if layout == "BCHW" else 1, stream_ptr()),
                   ctx.handle)
        self.called_forward = True
        return out
