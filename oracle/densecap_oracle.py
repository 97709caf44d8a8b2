"""CPU oracle for the DenseCap inference hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement of the reference algorithm (jcjohnson/densecap @ 7c32170,
Lua/Torch7).  It is the *checker* for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import it.  Nothing under ``densecap_b200/`` imports it and the product path has no
CPU fallback.

Every function cites the reference file:line it restates (paths relative to
/root/reference).  Box/index arithmetic is numpy fp32 with the reference's operation
order (one rounding per Torch tensor op, no FMA); dense contractions (conv / linear)
use torch-CPU fp32 (oneDNN / MKL) because the reference's own summation order
(THNN im2col + BLAS sgemm) is unspecified.

PARITY PINNING:
  * pinned by the reference's own known-answer tests (see tests/test_oracle_golden.py):
      nms (test/nms_test.lua:9-95), BoxToAffine (test/BoxToAffine_test.lua:14-44),
      ApplyBoxTransform (test/ApplyBoxTransform_test.lua:12-35), anchor/transform
      composition + row order (test/MakeBoxes_test.lua:47-234 through the equivalence
      stated by test/MakeAnchors_test.lua:17-44), ReshapeBoxFeatures strides
      (test/ReshapeBoxFeatures_test.lua:33-58), Apply/Invert round trip
      (test/InvertBoxTransform_test.lua:15-58), decodeSequence and getTarget
      (test/LanguageModel_test.lua:46-70,135-160), BoxSamplerHelper gather / scatter protocol
      (test/BoxSamplerHelper_test.lua), BoxRegressionCriterion numeric gradient check
      (test/BoxRegressionCriterion_test.lua), box conversion round trip (test/box_conversion_test.lua).
  * self-checked where the reference has no numbers: the training-time backward
    (localization_layer_backward_train) against central finite differences of this file's own
    float64 forward (7 digits at 720x600 / 256 ROIs).
  * "parity unpinned" (third-party arithmetic that is NOT under /root/reference and
    has no numeric known-answer test there; restated from the published algorithm):
      stnbhwd AffineGridGeneratorBHWD / BilinearSamplerBHWD  (sampler_* below),
      torch-rnn nn.LSTM gate order / layout                   (lstm_step below),
      torch/nn SpatialConvolution / SpatialMaxPooling(ceil) / Linear numerics.
    The reference itself (Lua/Torch7) cannot run in this environment (no th/luajit).
"""
from __future__ import annotations

import math
import numpy as np

# Working precision.  float32 = the reference's arithmetic (default, what parity is judged against).
# `with precision(np.float64):` re-runs the SAME restatement in float64 ("gold") so that tests can measure the
# fp32 summation-order noise floor of the reference path itself and budget tolerances against it.
_DT = [np.float32]


class precision:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        self.prev = _DT[0]
        _DT[0] = self.dtype
        return self

    def __exit__(self, *a):
        _DT[0] = self.prev


def f32(v):
    return _DT[0](v)


def _f(x):
    return np.asarray(x, dtype=_DT[0])


# ----------------------------------------------------------------------------------------
# exp as the reference's CPU path sees it: glibc expf is correctly rounded in all but
# ~1e-8 of cases, which is what exp-in-double-then-round gives.  The CUDA kernels use the
# same recipe ((float)exp((double)x)) so both sides agree bit for bit.
# ----------------------------------------------------------------------------------------
def exp_f32(x):
    return np.exp(np.asarray(x, dtype=np.float64)).astype(_DT[0])


# ----------------------------------------------------------------------------------------
# densecap/modules/MakeAnchors.lua:40-67
# ----------------------------------------------------------------------------------------
def make_anchors(N, H, W, x0, y0, sx, sy, anchors):
    """anchors: (2,k) widths/heights.  Returns (N,4k,H,W) fp32, channel 4a+{0:xc,1:yc,2:w,3:h}."""
    anchors = _f(anchors)
    k = anchors.shape[1]
    xc = np.arange(W, dtype=_DT[0]) * f32(sx) + f32(x0)      # :mul(sx):add(x0)
    yc = np.arange(H, dtype=_DT[0]) * f32(sy) + f32(y0)
    out = np.empty((N, k, 4, H, W), _DT[0])
    out[:, :, 0] = xc[None, None, None, :]
    out[:, :, 1] = yc[None, None, :, None]
    out[:, :, 2] = anchors[0][None, :, None, None]
    out[:, :, 3] = anchors[1][None, :, None, None]
    return out.reshape(N, 4 * k, H, W)


# ----------------------------------------------------------------------------------------
# densecap/modules/ReshapeBoxFeatures.lua:24-33   (row r = (a*H + y)*W + x)
# ----------------------------------------------------------------------------------------
def reshape_box_features(x, k):
    N, kD, H, W = x.shape
    D = kD // k
    return np.ascontiguousarray(x.reshape(N, k, D, H, W).transpose(0, 1, 3, 4, 2)).reshape(N, k * H * W, D)


# densecap/modules/ReshapeBoxFeatures.lua:36-44
def reshape_box_features_backward(grad_out, k, H, W):
    N, _, D = grad_out.shape
    return np.ascontiguousarray(grad_out.reshape(N, k, H, W, D).transpose(0, 1, 4, 2, 3)).reshape(N, k * D, H, W)


# ----------------------------------------------------------------------------------------
# densecap/modules/ApplyBoxTransform.lua:63-90
# ----------------------------------------------------------------------------------------
def apply_box_transform(boxes, trans):
    b = _f(boxes).reshape(-1, 4)
    t = _f(trans).reshape(-1, 4)
    out = np.empty_like(b)
    out[:, 0] = t[:, 0] * b[:, 2] + b[:, 0]          # cmul(tx, wa):add(xa)  (two roundings)
    out[:, 1] = t[:, 1] * b[:, 3] + b[:, 1]
    out[:, 2] = exp_f32(t[:, 2]) * b[:, 2]           # exp(tw):cmul(wa)
    out[:, 3] = exp_f32(t[:, 3]) * b[:, 3]
    return out.reshape(np.shape(boxes))


# densecap/modules/ApplyBoxTransform.lua:93-128
def apply_box_transform_backward(boxes, trans, out, grad_out):
    b = _f(boxes).reshape(-1, 4); t = _f(trans).reshape(-1, 4)
    o = _f(out).reshape(-1, 4); g = _f(grad_out).reshape(-1, 4)
    gb = np.zeros_like(b); gt = np.zeros_like(t)
    gt[:, 0] = g[:, 0] * b[:, 2]
    gt[:, 1] = g[:, 1] * b[:, 3]
    gt[:, 2] = o[:, 2] * g[:, 2]
    gt[:, 3] = o[:, 3] * g[:, 3]
    gb[:, 0] = g[:, 0]
    gb[:, 1] = g[:, 1]
    gb[:, 2] = exp_f32(t[:, 2]) * g[:, 2] + t[:, 0] * g[:, 0]
    gb[:, 3] = exp_f32(t[:, 3]) * g[:, 3] + t[:, 1] * g[:, 1]
    return gb.reshape(np.shape(boxes)), gt.reshape(np.shape(trans))


# ----------------------------------------------------------------------------------------
# densecap/modules/InvertBoxTransform.lua:36-60
# ----------------------------------------------------------------------------------------
def invert_box_transform(anchor_boxes, target_boxes):
    a = _f(anchor_boxes); t = _f(target_boxes)
    out = np.empty_like(t)
    out[:, 0] = (t[:, 0] - a[:, 0]) / a[:, 2]
    out[:, 1] = (t[:, 1] - a[:, 1]) / a[:, 3]
    out[:, 2] = np.log((t[:, 2] / a[:, 2]).astype(np.float64)).astype(_DT[0])
    out[:, 3] = np.log((t[:, 3] / a[:, 3]).astype(np.float64)).astype(_DT[0])
    return out


# ----------------------------------------------------------------------------------------
# densecap/box_utils.lua:270-298   x1 = xc - (w-1)/2 ; x2 = xc + (w-1)/2
# ----------------------------------------------------------------------------------------
def xcycwh_to_x1y1x2y2(boxes):
    b = _f(boxes)
    xc, yc, w, h = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    ret = np.empty_like(b)
    half_w = (w + f32(-1)) / f32(2.0)
    half_h = (h + f32(-1)) / f32(2.0)
    ret[..., 0] = half_w * f32(-1) + xc
    ret[..., 2] = half_w + xc
    ret[..., 1] = half_h * f32(-1) + yc
    ret[..., 3] = half_h + yc
    return ret


# densecap/box_utils.lua:382-410   xc = (x0+x1)/2 ; w = x1 - x0   (NOT the inverse of the above)
def x1y1x2y2_to_xcycwh(boxes):
    b = _f(boxes)
    x0, y0, x1, y1 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    out = np.empty_like(b)
    out[..., 0] = (x0 + x1) / f32(2.0)
    out[..., 1] = (y0 + y1) / f32(2.0)
    out[..., 2] = x1 - x0
    out[..., 3] = y1 - y0
    return out


# densecap/box_utils.lua:311-380 (xywh helpers used by run_model.lua:79)
def x1y1x2y2_to_xywh(boxes):
    b = _f(boxes)
    out = np.empty_like(b)
    out[..., 0] = b[..., 0]
    out[..., 1] = b[..., 1]
    out[..., 2] = (b[..., 2] - b[..., 0]) + f32(1)
    out[..., 3] = (b[..., 3] - b[..., 1]) + f32(1)
    return out


def xcycwh_to_xywh(boxes):
    return x1y1x2y2_to_xywh(xcycwh_to_x1y1x2y2(boxes))


# ----------------------------------------------------------------------------------------
# densecap/box_utils.lua:486-523  clip_boxes(boxes, bounds, 'xcycwh')
# ----------------------------------------------------------------------------------------
def clip_boxes(boxes, x_min, y_min, x_max, y_max, fmt="xcycwh"):
    shape = np.shape(boxes)
    if fmt == "xcycwh":
        c = xcycwh_to_x1y1x2y2(boxes).reshape(-1, 4)
    elif fmt == "x1y1x2y2":
        c = _f(boxes).reshape(-1, 4).copy()
    else:
        raise ValueError('Unrecognized box format "%s"' % fmt)
    c[:, 0] = np.clip(c[:, 0], f32(x_min), f32(x_max - 1))
    c[:, 1] = np.clip(c[:, 1], f32(y_min), f32(y_max - 1))
    c[:, 2] = np.clip(c[:, 2], f32(x_min + 1), f32(x_max))
    c[:, 3] = np.clip(c[:, 3], f32(y_min + 1), f32(y_max))
    valid = (c[:, 2] > c[:, 0]) & (c[:, 3] > c[:, 1])
    if fmt == "xcycwh":
        c = x1y1x2y2_to_xcycwh(c)
    return c.reshape(shape), valid.astype(np.uint8)


# ----------------------------------------------------------------------------------------
# densecap/box_utils.lua:154-256   greedy NMS; returns 1-based int64 indices, best first.
# Sort: reference uses TH's unstable ascending quicksort on CPU float (:185) and pops from
# the end; the oracle DEFINES the tie rule as a stable ascending sort, i.e. among equal
# scores the HIGHER index is picked first.
# ----------------------------------------------------------------------------------------
def nms(boxes_scores, overlap, max_boxes=None):
    bs = _f(boxes_scores)
    if bs.size == 0:
        return np.zeros((0,), np.int64)
    x1, y1, x2, y2, s = bs[:, 0], bs[:, 1], bs[:, 2], bs[:, 3], bs[:, 4]
    area = ((x2 - x1) + f32(1.0)) * ((y2 - y1) + f32(1.0))
    I = np.argsort(s, kind="stable")                     # ascending
    thr = f32(overlap)
    pick = []
    while (max_boxes is None or len(pick) < max_boxes) and I.size > 0:
        i = I[-1]
        pick.append(i)
        if I.size == 1:
            break
        I = I[:-1]
        xx1 = np.maximum(x1[I], x1[i]); xx2 = np.minimum(x2[I], x2[i])
        yy1 = np.maximum(y1[I], y1[i]); yy2 = np.minimum(y2[I], y2[i])
        w = np.maximum((xx2 - xx1) + f32(1.0), f32(0))
        h = np.maximum((yy2 - yy1) + f32(1.0), f32(0))
        inter = w * h
        union = (area[I] + area[i]) - inter
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / union
        I = I[iou <= thr]                                  # NaN -> dropped, as torch le()
    return np.asarray(pick, np.int64) + 1


# ----------------------------------------------------------------------------------------
# densecap/modules/BoxToAffine.lua:69-93      (H, W = IMAGE size)
# ----------------------------------------------------------------------------------------
def box_to_affine(boxes, H, W):
    b = _f(boxes)
    B = b.shape[0]
    th = np.zeros((B, 2, 3), _DT[0])
    th[:, 1, 2] = (b[:, 0] * f32(2) + f32(-1 - W)) / f32(W - 1)
    th[:, 0, 2] = (b[:, 1] * f32(2) + f32(-1 - H)) / f32(H - 1)
    th[:, 1, 1] = b[:, 2] / f32(W)
    th[:, 0, 0] = b[:, 3] / f32(H)
    return th


# densecap/modules/BoxToAffine.lua:96-118
def box_to_affine_backward(grad_theta, H, W):
    g = _f(grad_theta)
    gi = np.zeros((g.shape[0], 4), _DT[0])
    gi[:, 0] = g[:, 1, 2] * f32(2 / (W - 1))
    gi[:, 1] = g[:, 0, 2] * f32(2 / (H - 1))
    gi[:, 2] = g[:, 1, 1] * f32(1 / W)
    gi[:, 3] = g[:, 0, 0] * f32(1 / H)
    return gi


# ----------------------------------------------------------------------------------------
# [3P, parity unpinned] qassemoquab/stnbhwd AffineGridGeneratorBHWD(HH, WW)
#   call site densecap/modules/BilinearRoiPooling.lua:52 ; SURVEY.md Appendix B1.
#   baseGrid[i][j] = (y_i, x_j, 1), y_i = -1 + (i-1)/(HH-1)*2 (Lua double -> float tensor);
#   grid[b] = baseGrid . theta[b]^T  -> last dim (gy, gx).
# ----------------------------------------------------------------------------------------
def affine_grid(theta, HH, WW):
    th = _f(theta)
    ys = np.array([-1 + (i / (HH - 1)) * 2 for i in range(HH)], np.float64).astype(_DT[0])
    xs = np.array([-1 + (j / (WW - 1)) * 2 for j in range(WW)], np.float64).astype(_DT[0])
    Y = ys[None, :, None]; X = xs[None, None, :]
    gy = (th[:, 0, 0, None, None] * Y + th[:, 0, 1, None, None] * X) + th[:, 0, 2, None, None]
    gx = (th[:, 1, 0, None, None] * Y + th[:, 1, 1, None, None] * X) + th[:, 1, 2, None, None]
    return np.stack([gy, gx], axis=-1).astype(_DT[0])          # (B,HH,WW,2)


# ----------------------------------------------------------------------------------------
# [3P, parity unpinned] stnbhwd BilinearSamplerBHWD_updateOutput
#   call site densecap/modules/BatchBilinearSamplerBHWD.lua:104-122 ; SURVEY.md App. B2.
# ----------------------------------------------------------------------------------------
def bilinear_sample(feats_hwc, grids):
    """feats_hwc (H,W,C) fp32, grids (B,HH,WW,2) (y,x) in [-1,1] -> (B,HH,WW,C)."""
    feats = _f(feats_hwc)
    H, W, C = feats.shape
    g = _f(grids)
    gy, gx = g[..., 0], g[..., 1]
    xcoord = ((gx + f32(1)) * f32(W - 1)) / f32(2)
    ycoord = ((gy + f32(1)) * f32(H - 1)) / f32(2)
    xL = np.floor(xcoord); yT = np.floor(ycoord)
    wx = f32(1) - (xcoord - xL)
    wy = f32(1) - (ycoord - yT)
    xL = xL.astype(np.int64); yT = yT.astype(np.int64)

    def tap(y, x):
        ok = (x >= 0) & (x <= W - 1) & (y >= 0) & (y <= H - 1)
        v = feats[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)]     # (B,HH,WW,C)
        return np.where(ok[..., None], v, f32(0))

    TL = tap(yT, xL); TR = tap(yT, xL + 1); BL = tap(yT + 1, xL); BR = tap(yT + 1, xL + 1)
    wx_ = wx[..., None]; wy_ = wy[..., None]
    omx = f32(1) - wx_; omy = f32(1) - wy_
    out = (wx_ * wy_) * TL
    out = out + (omx * wy_) * TR
    out = out + (wx_ * omy) * BL
    out = out + (omx * omy) * BR
    return out.astype(_DT[0])


# stnbhwd BilinearSamplerBHWD_updateGradInput (App. B2), summed over the batch as in
# densecap/modules/BatchBilinearSamplerBHWD.lua:125-157.  Accumulates in float64 (gold).
def bilinear_sample_backward(feats_hwc, grids, grad_out):
    feats = _f(feats_hwc); H, W, C = feats.shape
    g = _f(grids); go = _f(grad_out)
    gy, gx = g[..., 0], g[..., 1]
    xcoord = ((gx + f32(1)) * f32(W - 1)) / f32(2)
    ycoord = ((gy + f32(1)) * f32(H - 1)) / f32(2)
    xL = np.floor(xcoord); yT = np.floor(ycoord)
    wx = (f32(1) - (xcoord - xL)).astype(np.float64); wy = (f32(1) - (ycoord - yT)).astype(np.float64)
    xL = xL.astype(np.int64); yT = yT.astype(np.int64)
    gfeat = np.zeros((H, W, C), np.float64)
    ggrid = np.zeros(g.shape, np.float64)
    go64 = go.astype(np.float64)
    dots = {}
    for name, (dy, dx, wgt) in {"TL": (0, 0, wx * wy), "TR": (0, 1, (1 - wx) * wy),
                                "BL": (1, 0, wx * (1 - wy)), "BR": (1, 1, (1 - wx) * (1 - wy))}.items():
        y = yT + dy; x = xL + dx
        ok = (x >= 0) & (x <= W - 1) & (y >= 0) & (y <= H - 1)
        yy = np.clip(y, 0, H - 1); xx = np.clip(x, 0, W - 1)
        contrib = np.where(ok[..., None], wgt[..., None] * go64, 0.0)
        np.add.at(gfeat, (yy.reshape(-1), xx.reshape(-1)), contrib.reshape(-1, C))
        v = np.where(ok[..., None], feats[yy, xx].astype(np.float64), 0.0)
        dots[name] = (v * go64).sum(-1)
    ggrid[..., 0] = (-wx * dots["TL"] + wx * dots["BL"] - (1 - wx) * dots["TR"] + (1 - wx) * dots["BR"]) * (H - 1) / 2
    ggrid[..., 1] = (-wy * dots["TL"] + wy * dots["TR"] - (1 - wy) * dots["BL"] + (1 - wy) * dots["BR"]) * (W - 1) / 2
    return gfeat.astype(_DT[0]), ggrid.astype(_DT[0])


# ----------------------------------------------------------------------------------------
# densecap/modules/BilinearRoiPooling.lua:35-67,84-91
#   feats (C,H',W') -> transpose HWC ; boxes (B,4) xcycwh in IMAGE coords ; -> (B,C,HH,WW)
# ----------------------------------------------------------------------------------------
def bilinear_roi_pooling(feats_chw, boxes, image_h, image_w, HH=7, WW=7, layout="BCHW"):
    feats_hwc = np.ascontiguousarray(_f(feats_chw).transpose(1, 2, 0))
    theta = box_to_affine(boxes, image_h, image_w)
    grids = affine_grid(theta, HH, WW)
    out = bilinear_sample(feats_hwc, grids)                       # (B,HH,WW,C)
    if layout == "BHWC":
        return out
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2))


# ----------------------------------------------------------------------------------------
# densecap/modules/BoxIoU.lua:40-73  (area = w*h, corners from (w-1)/2, no +1)
# ----------------------------------------------------------------------------------------
def box_iou(box1, box2):
    b1 = _f(box1); b2 = _f(box2)                     # (N,B1,4), (N,B2,4)
    a1 = (b1[..., 2] * b1[..., 3])[:, :, None]
    a2 = (b2[..., 2] * b2[..., 3])[:, None, :]
    l1 = xcycwh_to_x1y1x2y2(b1)[:, :, None, :]
    l2 = xcycwh_to_x1y1x2y2(b2)[:, None, :, :]
    x0 = np.maximum(l1[..., 0], l2[..., 0]); y0 = np.maximum(l1[..., 1], l2[..., 1])
    x1 = np.minimum(l1[..., 2], l2[..., 2]); y1 = np.minimum(l1[..., 3], l2[..., 3])
    w = np.maximum(x1 - x0, f32(0)); h = np.maximum(y1 - y0, f32(0))
    inter = w * h
    with np.errstate(divide="ignore", invalid="ignore"):
        out = (f32(1) / ((a1 - inter) + a2)) * inter          # add(a1,-1,inter):add(a2):pow(-1):cmul(inter)
    return out.astype(_DT[0])


# ========================================================================================
# Dense layers (torch CPU fp32).  [3P, parity unpinned] torch/nn definitions, App. B4.
# ========================================================================================
def _torch():
    import torch
    return torch


def conv2d(x, w, b, pad):
    torch = _torch()
    import torch.nn.functional as F
    y = F.conv2d(torch.from_numpy(_f(x)), torch.from_numpy(_f(w)),
                 None if b is None else torch.from_numpy(_f(b)), stride=1, padding=pad)
    return y.numpy()


def max_pool2x2(x, ceil_mode=True):
    torch = _torch()
    import torch.nn.functional as F
    return F.max_pool2d(torch.from_numpy(_f(x)), 2, 2, 0, ceil_mode=ceil_mode).numpy()


def linear(x, w, b):
    torch = _torch()
    y = torch.from_numpy(_f(x)) @ torch.from_numpy(_f(w)).t()
    if b is not None:
        y = y + torch.from_numpy(_f(b))
    return y.numpy()


def relu(x):
    return np.maximum(x, f32(0))


VGG16_LAYERS = ["conv1_1", "conv1_2", "P", "conv2_1", "conv2_2", "P",
                "conv3_1", "conv3_2", "conv3_3", "P", "conv4_1", "conv4_2", "conv4_3", "P",
                "conv5_1", "conv5_2", "conv5_3"]


# densecap/DenseCapModel.lua:61-76 (conv_net1 = layers 1-10, conv_net2 = 11-30; pool5 excluded)
def vgg_trunk(img, weights, pool_ceil=True, taps=None):
    x = _f(img)
    for name in VGG16_LAYERS:
        if name == "P":
            x = max_pool2x2(x, pool_ceil)
        else:
            x = relu(conv2d(x, weights[name + ".weight"], weights[name + ".bias"], 1))
        if taps is not None and name != "P":
            taps[name] = x
    return x


DEFAULT_ANCHORS = np.array([[45, 90], [90, 45], [64, 64], [90, 180], [180, 90], [128, 128],
                            [181, 362], [362, 181], [256, 256], [362, 724], [724, 362], [512, 512]],
                           np.float32).T.copy()          # (2,k)   LocalizationLayer.lua:613-618
FIELD_CENTERS = (8.5, 8.5, 16.0, 16.0)                   # net_utils.lua:106-140 on VGG-16 1..30


# densecap/LocalizationLayer.lua:609-690  build_rpn -> {boxes, anchors, trans, scores}
def rpn_forward(feats, weights, anchors=DEFAULT_ANCHORS, field_centers=FIELD_CENTERS):
    k = anchors.shape[1]
    N, C, H, W = feats.shape
    hid = relu(conv2d(feats, weights["rpn.conv.weight"], weights["rpn.conv.bias"], 1))
    box_out = conv2d(hid, weights["rpn.box.weight"], weights["rpn.box.bias"], 0)       # (N,4k,H,W)
    score_out = conv2d(hid, weights["rpn.score.weight"], weights["rpn.score.bias"], 0)  # (N,2k,H,W)
    x0, y0, sx, sy = field_centers
    anc = reshape_box_features(make_anchors(N, H, W, x0, y0, sx, sy, anchors), k)
    trans = reshape_box_features(box_out, k)
    boxes = apply_box_transform(anc, trans)
    scores = reshape_box_features(score_out, k)
    return boxes, anc, trans, scores, hid


# densecap/LocalizationLayer.lua:250-363 with rpn outputs given (post-processing only)
def rpn_postprocess(rpn_boxes, rpn_anchors, rpn_trans, rpn_scores, image_h, image_w,
                    nms_thresh=0.7, max_proposals=300, clip=True):
    boxes, anchors, trans, scores = rpn_boxes, rpn_anchors, rpn_trans, rpn_scores
    valid = None
    if clip:
        boxes, valid = clip_boxes(boxes, 1, 1, image_w, image_h, "xcycwh")
        m = valid.astype(bool)
        boxes = boxes[:, m]; anchors = anchors[:, m]; trans = trans[:, m]; scores = scores[:, m]
    x1y1x2y2 = xcycwh_to_x1y1x2y2(boxes)
    e = exp_f32(scores)
    pos, neg = e[0, :, 0], e[0, :, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        prob = (f32(1) / (pos + neg)) * pos                 # (pos+neg):pow(-1):cmul(pos)
    bs = np.concatenate([x1y1x2y2[0], prob[:, None]], axis=1)
    idx = nms(bs, nms_thresh, None if max_proposals == -1 else max_proposals)
    i0 = idx - 1
    return {"boxes": boxes[0, i0], "anchors": anchors[0, i0], "trans": trans[0, i0],
            "scores": scores[0, i0], "prob": prob[i0], "idx": idx, "valid": valid,
            "all_prob": prob, "all_boxes_x1y1x2y2": x1y1x2y2[0]}


def localization_layer_forward_test(feats, weights, image_h, image_w, nms_thresh=0.7,
                                    max_proposals=300, clip=True, anchors=DEFAULT_ANCHORS,
                                    field_centers=FIELD_CENTERS, HH=7, WW=7):
    boxes, anc, trans, scores, _ = rpn_forward(feats, weights, anchors, field_centers)
    pp = rpn_postprocess(boxes, anc, trans, scores, image_h, image_w, nms_thresh, max_proposals, clip)
    roi = bilinear_roi_pooling(feats[0], pp["boxes"], image_h, image_w, HH, WW)
    pp["roi_features"] = roi
    return pp


# densecap/DenseCapModel.lua:87-100,127-162  recognition head (test time)
def recog_head(roi_features, roi_boxes, weights):
    B = roi_features.shape[0]
    x = roi_features.reshape(B, -1)                                   # View: (c,y,x) flatten
    x = relu(linear(x, weights["fc6.weight"], weights["fc6.bias"]))
    codes = relu(linear(x, weights["fc7.weight"], weights["fc7.bias"]))
    obj = linear(codes, weights["obj.weight"], weights["obj.bias"])              # (B,1)
    trans = linear(codes, weights["boxreg.weight"], weights["boxreg.bias"])      # (B,4)
    final_boxes = apply_box_transform(roi_boxes, trans)
    return codes, obj, trans, final_boxes


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(_DT[0])


# [3P, parity unpinned] jcjohnson/torch-rnn nn.LSTM step, SURVEY.md App. B3:
#   weight ((D+H),4H) rows [Wx;Wh], gate blocks [i|f|o|g]
def lstm_step(x, h, c, W, b):
    D = x.shape[1]; H = h.shape[1]
    a = linear(x, W[:D].T, None) + linear(h, W[D:].T, None) + _f(b)[None, :]
    i = sigmoid(a[:, 0:H]); f = sigmoid(a[:, H:2 * H]); o = sigmoid(a[:, 2 * H:3 * H])
    g = np.tanh(a[:, 3 * H:4 * H].astype(np.float64)).astype(_DT[0])
    c2 = f * c + i * g
    h2 = o * np.tanh(c2.astype(np.float64)).astype(_DT[0])
    return h2, c2


# densecap/LanguageModel.lua:293-348  LM:sample (greedy argmax), tokens 1-based
def lm_sample(codes, weights, seq_length, return_logits=False):
    N = codes.shape[0]
    Wl = weights["lm.lstm.weight"]; bl = weights["lm.lstm.bias"]
    H = Wl.shape[1] // 4
    V1 = weights["lm.vocab.weight"].shape[0]            # V+1
    START = V1                                           # START = END = V+1  (:32-33)
    enc = relu(linear(codes, weights["lm.encoder.weight"], weights["lm.encoder.bias"]))
    h = np.zeros((N, H), _DT[0]); c = np.zeros((N, H), _DT[0])
    h, c = lstm_step(enc, h, c, Wl, bl)                  # first step: image vector, output ignored (:312-313)
    seq = np.zeros((N, seq_length), np.int64)
    logits_all = []
    for t in range(seq_length):
        words = np.full((N,), START, np.int64) if t == 0 else seq[:, t - 1]
        x = _f(weights["lm.lookup.weight"])[words - 1]
        h, c = lstm_step(x, h, c, Wl, bl)
        logits = linear(h, weights["lm.vocab.weight"], weights["lm.vocab.bias"])
        seq[:, t] = np.argmax(logits, axis=1) + 1        # torch.max: first maximal index
        if return_logits:
            logits_all.append(logits)
    if return_logits:
        return seq, np.stack(logits_all, axis=1)
    return seq


# densecap/LanguageModel.lua:86-103
def decode_sequence(seq, idx_to_token, end_token):
    caps = []
    for row in np.asarray(seq):
        words = []
        for idx in row:
            if idx == end_token or idx == 0:
                break
            words.append(idx_to_token[int(idx)])
        caps.append(" ".join(words))
    return caps


# densecap/DenseCapModel.lua:242-282 (test time)  + forward_test :319-327
def densecap_forward_test(img, weights, cfg, return_intermediates=False):
    """img (1,3,H,W) fp32.  cfg: dict(rpn_nms_thresh, final_nms_thresh, num_proposals, seq_length, pool_ceil)."""
    H, W = img.shape[2], img.shape[3]
    feats = vgg_trunk(img, weights, cfg.get("pool_ceil", True))
    ll = localization_layer_forward_test(feats, weights, H, W, cfg.get("rpn_nms_thresh", 0.7),
                                         cfg.get("num_proposals", 1000), True)
    codes, obj, trans, final_boxes = recog_head(ll["roi_features"], ll["boxes"], weights)
    seq = lm_sample(codes, weights, cfg["seq_length"])
    out = {"objectness_scores": obj, "roi_boxes": ll["boxes"], "final_box_trans": trans,
           "final_boxes": final_boxes, "seq": seq}
    thr = cfg.get("final_nms_thresh", 0.3)
    if thr > 0:
        bs = np.concatenate([xcycwh_to_x1y1x2y2(final_boxes), obj[:, :1]], axis=1)
        idx = nms(bs, thr)
        out["final_idx"] = idx
        out["final_boxes"] = final_boxes[idx - 1]
        out["objectness_scores"] = obj[idx - 1]
        out["seq"] = seq[idx - 1]
    if return_intermediates:
        out["feats"] = feats; out["ll"] = ll; out["codes"] = codes
        out["pre_nms"] = {"final_boxes": final_boxes, "obj": obj, "seq": seq}
    return out


# ========================================================================================
# Training-time forward of the LocalizationLayer (BASELINE config 5)
# ========================================================================================

# densecap/modules/BoxSampler.lua:64-167.  Sampling: torch.multinomial is irreproducible, so -- exactly like the
# reference's own debug hook (debug_pos_sample_idx / debug_neg_sample_idx, BoxSampler.lua:154-159) -- indices INTO the
# positive / negative candidate lists are injected; the default injection is "the first P / first M candidates".
def box_sampler(input_boxes, target_boxes, low_thresh=0.5, high_thresh=0.7, batch_size=256, bounds=None,
                pos_sample_idx=None, neg_sample_idx=None):
    """input_boxes (1,B1,4), target_boxes (1,B2,4) xcycwh.  Returns 1-based (pos_input_idx, pos_target_idx, neg_input_idx)
    plus the masks (for tests)."""
    ious = box_iou(input_boxes, target_boxes)[0]                      # (B1,B2)
    B1 = ious.shape[0]
    input_max_iou = ious.max(axis=1)
    input_idx = ious.argmax(axis=1)                                   # first maximal target per input (torch.max)
    target_idx = ious.argmax(axis=0)                                  # best input per target
    pos_mask = input_max_iou > f32(high_thresh)
    neg_mask = input_max_iou < f32(low_thresh)
    if bounds is not None:                                            # BoxSampler.lua:82-98
        c = xcycwh_to_x1y1x2y2(input_boxes)[0]
        oob = (c[:, 0] < f32(bounds["x_min"])) | (c[:, 1] < f32(bounds["y_min"])) | \
              (c[:, 2] > f32(bounds["x_max"])) | (c[:, 3] > f32(bounds["y_max"]))
        pos_mask = pos_mask & ~oob
        neg_mask = neg_mask & ~oob
    pos_mask = pos_mask.copy(); neg_mask = neg_mask.copy()
    pos_mask[target_idx] = True                                       # :104-105 force the best match of every GT box
    neg_mask[target_idx] = False
    if neg_mask.sum() == 0:                                           # :111-128
        neg_mask = ~pos_mask
    pos_nz = np.nonzero(pos_mask)[0]
    neg_nz = np.nonzero(neg_mask)[0]
    num_pos = int(min(batch_size // 2, pos_nz.size))
    num_neg = int(batch_size - num_pos)
    if pos_sample_idx is None:
        pos_sample_idx = np.arange(num_pos)
    if neg_sample_idx is None:
        neg_sample_idx = np.arange(num_neg) % max(neg_nz.size, 1)     # with replacement when there are too few
    pos_input_idx = pos_nz[np.asarray(pos_sample_idx, np.int64)]
    neg_input_idx = neg_nz[np.asarray(neg_sample_idx, np.int64)]
    pos_target_idx = input_idx[pos_input_idx]
    return pos_input_idx + 1, pos_target_idx + 1, neg_input_idx + 1, pos_mask, neg_mask


# densecap/modules/BoxSamplerHelper.lua:68-129 (gather) / :142-182 (scatter of gradients, indexCopy semantics)
def box_sampler_helper_forward(input_data, target_data, pos_input_idx, pos_target_idx, neg_input_idx):
    pos = [d[0][pos_input_idx - 1] for d in input_data]
    tgt = [d[0][pos_target_idx - 1] for d in target_data]
    neg = [d[0][neg_input_idx - 1] for d in input_data]
    return pos, tgt, neg


def box_sampler_helper_backward(input_shapes, pos_input_idx, neg_input_idx, grad_pos, grad_neg):
    out = []
    for i, shp in enumerate(input_shapes):
        g = np.zeros(shp, _DT[0])
        if grad_pos[i] is not None:
            g[0][pos_input_idx - 1] = grad_pos[i]          # indexCopy: later duplicates overwrite
        if grad_neg[i] is not None:
            g[0][neg_input_idx - 1] = grad_neg[i]
        out.append(g)
    return out


# nn.LogSoftMax + nn.ClassNLLCriterion (sizeAverage) = densecap/modules/OurCrossEntropyCriterion.lua:15-32 ; target 1-based
def cross_entropy(scores, target):
    s = _f(scores).astype(np.float64)
    m = s.max(axis=1, keepdims=True)
    logp = s - m - np.log(np.exp(s - m).sum(axis=1, keepdims=True))
    n = s.shape[0]
    loss = -logp[np.arange(n), np.asarray(target) - 1].sum() / max(n, 1)
    grad = np.exp(logp)
    grad[np.arange(n), np.asarray(target) - 1] -= 1.0
    return f32(loss), (grad / max(n, 1)).astype(_DT[0])


# [3P] nn.SmoothL1Criterion (sizeAverage), SURVEY.md App. B4
def smooth_l1(x, t):
    d = _f(x).astype(np.float64) - _f(t).astype(np.float64)
    n = d.size
    a = np.abs(d)
    loss = np.where(a < 1, 0.5 * d * d, a - 0.5).sum() / max(n, 1)
    grad = np.clip(d, -1, 1) / max(n, 1)
    return f32(loss), grad.astype(_DT[0])


# densecap/LocalizationLayer.lua:383-526  _forward_train (losses in stats.losses)
def localization_layer_forward_train(feats, weights, gt_boxes, image_h, image_w, mid_box_reg_weight=0.05,
                                     mid_objectness_weight=0.1, box_reg_decay=0.0, sampler_low=0.5, sampler_high=0.7,
                                     batch_size=256, anchors=DEFAULT_ANCHORS, field_centers=FIELD_CENTERS, HH=7, WW=7,
                                     pos_sample_idx=None, neg_sample_idx=None):
    """feats (1,C,H',W'), gt_boxes (1,B1,4).  Default sampling = first P positives / first M negatives; pos_sample_idx /
    neg_sample_idx inject other draws (0-based indices into the candidate lists), like BoxSampler.lua:154-159."""
    boxes, anc, trans, scores, _ = rpn_forward(feats, weights, anchors, field_centers)
    bounds = {"x_min": 1, "y_min": 1, "x_max": image_w, "y_max": image_h}          # :414-422
    pi, pt, ni, pos_mask, neg_mask = box_sampler(boxes, gt_boxes, sampler_low, sampler_high, batch_size, bounds,
                                                 pos_sample_idx=pos_sample_idx, neg_sample_idx=neg_sample_idx)
    pos, tgt, neg = box_sampler_helper_forward([boxes, anc, trans, scores], [gt_boxes], pi, pt, ni)
    pos_boxes, pos_anchors, pos_trans, pos_scores = pos
    pos_target_boxes = tgt[0]
    neg_boxes, neg_scores = neg[0], neg[3]
    num_pos, num_neg = pos_boxes.shape[0], neg_scores.shape[0]
    roi_boxes = np.concatenate([pos_boxes, neg_boxes], 0)                            # :444-446
    roi_features = bilinear_roi_pooling(feats[0], roi_boxes, image_h, image_w, HH, WW)
    lp, gp = cross_entropy(pos_scores, np.full((num_pos,), 1))                       # :464-467 (1 = positive)
    ln, gn = cross_entropy(neg_scores, np.full((num_neg,), 2))
    targets = invert_box_transform(pos_anchors, pos_target_boxes)                    # :474-477
    pos_trans = pos_trans.copy()
    big = np.abs(targets).max(axis=1) > 10                                           # :481-489 mask hack
    pos_trans[big] = 0
    targets = targets.copy(); targets[big] = 0
    lb, gb = smooth_l1(pos_trans, targets)
    decay = f32(0.5 * box_reg_decay * float(np.sum(_f(trans).astype(np.float64) ** 2)))   # RegularizeLayer.lua:11-16
    losses = {"obj_loss_pos": f32(mid_objectness_weight) * lp, "obj_loss_neg": f32(mid_objectness_weight) * ln,
              "box_reg_loss": f32(mid_box_reg_weight) * lb, "box_decay_loss": decay}
    losses["total_loss"] = f32(sum(float(v) for v in losses.values()))
    return {"roi_features": roi_features, "roi_boxes": roi_boxes, "pos_target_boxes": pos_target_boxes,
            "pos_input_idx": pi, "pos_target_idx": pt, "neg_input_idx": ni, "num_pos": num_pos, "num_neg": num_neg,
            "losses": losses, "pos_mask": pos_mask, "neg_mask": neg_mask, "masked": int(big.sum()),
            "grad_pos_scores": gp, "grad_neg_scores": gn, "grad_pos_trans": gb}


# densecap/modules/BilinearRoiPooling.lua:94-104 -> self.net:backward: sampler backward (w.r.t. features and grid),
# stnbhwd AffineGridGeneratorBHWD backward (gradTheta[b] = gradGrid[b]^T . baseGrid, App. B1), BoxToAffine backward.
def bilinear_roi_pooling_backward(feats_chw, boxes, image_h, image_w, grad_out_bchw, HH=7, WW=7):
    feats_hwc = np.ascontiguousarray(_f(feats_chw).transpose(1, 2, 0))
    theta = box_to_affine(boxes, image_h, image_w)
    grids = affine_grid(theta, HH, WW)
    go = np.ascontiguousarray(_f(grad_out_bchw).transpose(0, 2, 3, 1))           # BCHW -> BHWC
    gfeat_hwc, ggrid = bilinear_sample_backward(feats_hwc, grids, go)
    ys = np.array([-1 + (i / (HH - 1)) * 2 for i in range(HH)], np.float64).astype(_DT[0])
    xs = np.array([-1 + (j / (WW - 1)) * 2 for j in range(WW)], np.float64).astype(_DT[0])
    base = np.stack([np.broadcast_to(ys[:, None], (HH, WW)), np.broadcast_to(xs[None, :], (HH, WW)),
                     np.ones((HH, WW), _DT[0])], axis=-1).astype(np.float64)       # (HH,WW,3) = (y, x, 1)
    gtheta = np.einsum("bijd,ijk->bdk", ggrid.astype(np.float64), base).astype(_DT[0])   # (B,2,3)
    gboxes = box_to_affine_backward(gtheta, image_h, image_w)
    return np.ascontiguousarray(gfeat_hwc.transpose(2, 0, 1)), gboxes


# densecap/LocalizationLayer.lua:529-605  updateGradInput (+ accGradParameters of the RPN convolutions through
# nn.Module:backward).  `fwd` is the dict returned by localization_layer_forward_train for the same inputs.
# Dense parts (conv backward) use torch autograd on the plain fp32 torch definition of the RPN (App. B4).
def localization_layer_backward_train(feats, weights, gt_boxes, image_h, image_w, fwd, grad_roi_features, grad_roi_boxes,
                                      mid_box_reg_weight=0.05, mid_objectness_weight=0.1, box_reg_decay=0.0,
                                      anchors=DEFAULT_ANCHORS, field_centers=FIELD_CENTERS, HH=7, WW=7):
    torch = _torch()
    k = anchors.shape[1]
    N, C, H, W = feats.shape
    A = k * H * W
    P, M = fwd["num_pos"], fwd["num_neg"]
    pi, ni = fwd["pos_input_idx"], fwd["neg_input_idx"]
    # criterion gradients, scaled in place by the loss weights (:553-570)
    grad_pos_trans = fwd["grad_pos_trans"] * f32(mid_box_reg_weight)
    grad_pos_scores = fwd["grad_pos_scores"] * f32(mid_objectness_weight)
    grad_neg_scores = fwd["grad_neg_scores"] * f32(mid_objectness_weight)
    # RoI pooling backward (:573-580)
    gfeat_roi, gboxes_roi = bilinear_roi_pooling_backward(feats[0], fwd["roi_boxes"], image_h, image_w, grad_roi_features, HH, WW)
    grad_roi_boxes = _f(grad_roi_boxes) + gboxes_roi
    # sampler backward (:583-596): indexCopy scatter; anchors receive no gradient
    g_boxes, g_anc, g_trans, g_scores = box_sampler_helper_backward(
        [(1, A, 4), (1, A, 4), (1, A, 4), (1, A, 2)], pi, ni,
        [grad_roi_boxes[:P], None, grad_pos_trans, grad_pos_scores], [grad_roi_boxes[P:P + M], None, None, grad_neg_scores])
    # RPN backward (:599-602): ApplyBoxTransform backward + Identity branch, ReshapeBoxFeatures backward, RegularizeLayer backward
    boxes, anc, trans, scores, hid = rpn_forward(feats, weights, anchors, field_centers)
    _, gt_from_apply = apply_box_transform_backward(anc, trans, boxes, g_boxes)
    g_trans_total = g_trans + gt_from_apply
    g_box_out = reshape_box_features_backward(g_trans_total, k, H, W)
    box_out = reshape_box_features_backward(trans, k, H, W)                   # the box conv output itself
    g_box_out = (box_out * f32(box_reg_decay) + g_box_out).astype(_DT[0])     # RegularizeLayer.lua:18-22 (gradMul = 1)
    g_score_out = reshape_box_features_backward(g_scores, k, H, W)
    tf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).requires_grad_(True)
    x = tf(feats); wc = tf(weights["rpn.conv.weight"]); bc = tf(weights["rpn.conv.bias"])
    wb = tf(weights["rpn.box.weight"]); bb = tf(weights["rpn.box.bias"])
    ws = tf(weights["rpn.score.weight"]); bs = tf(weights["rpn.score.bias"])
    F = torch.nn.functional
    h = F.relu(F.conv2d(x, wc, bc, padding=1))
    bo = F.conv2d(h, wb, bb); so = F.conv2d(h, ws, bs)
    grads = torch.autograd.grad([bo, so], [x, wc, bc, wb, bb, ws, bs],
                                [torch.from_numpy(np.ascontiguousarray(g_box_out, dtype=np.float32)),
                                 torch.from_numpy(np.ascontiguousarray(g_score_out, dtype=np.float32))])
    gx = grads[0].numpy()
    grad_feats = (gx + gfeat_roi[None]).astype(_DT[0])
    names = ["rpn.conv.weight", "rpn.conv.bias", "rpn.box.weight", "rpn.box.bias", "rpn.score.weight", "rpn.score.bias"]
    return {"grad_feats": grad_feats, "grad_feats_roi": gfeat_roi, "grad_feats_rpn": gx, "grad_roi_boxes_pool": gboxes_roi,
            "grad_box_out": g_box_out, "grad_score_out": g_score_out,
            "grad_params": {n: g.numpy() for n, g in zip(names, grads[1:])}}


# run_model.lua:66-74 (and webcam/daemon.lua:66-78) without image.scale: image.load gives float RGB CHW in [0,1]
# (byte / 255); index {3,2,1} -> BGR; mul(255); add(-1, vgg_mean).  Input: uint8 (H,W,3) RGB.  Output (1,3,H,W) fp32.
VGG_MEAN_BGR = np.array([103.939, 116.779, 123.68], np.float32)
def preprocess_u8(img_u8_hwc):
    img = np.asarray(img_u8_hwc, np.uint8)
    x = img.astype(np.float32) / f32(255)                       # image.load
    x = np.ascontiguousarray(x.transpose(2, 0, 1)[::-1])        # CHW, BGR
    x = x * f32(255)
    x = x + (f32(-1) * VGG_MEAN_BGR)[:, None, None]
    return x[None].astype(np.float32)


# ----------------------------------------------------------------------------------------
# [3P-recalled, parity unpinned] torch/image `image.scale(src, size)` as called by run_model.lua:69 and
# webcam/daemon.lua:70 (the `image` rock is not under /root/reference; restated from its published source
# image/init.lua `scale` + generic/image.c `scaleBilinear` / `Main_scaleLinear_rowcol`):
#   * a NUMBER size rescales so that the LARGER side equals size: height = iheight*size/imax, width = iwidth*size/imax,
#     the tensor allocation truncates the fractions;
#   * bilinear mode = two separable 1-D passes in float (width first into a (C, srcH, dstW) temporary, then height):
#       upscaling   (dst_len > src_len): scale = (src_len-1)/(dst_len-1); dst[di] = (1-f)*src[i] + f*src[i+1] with
#                   i = floor(di*scale), f = di*scale - i; the last element is copied;
#       downscaling (dst_len < src_len): scale = src_len/dst_len; area average of the source span [di*scale, (di+1)*scale)
#                   with fractional end weights, acc / n;
#       equal length: copy.
# ----------------------------------------------------------------------------------------
def image_scaled_size(h, w, size):
    imax = max(h, w)
    return int(h * size / imax), int(w * size / imax)


def _scale_linear_1d(src, dst_len):
    """src (..., src_len) float32 -> (..., dst_len), the rowcol routine applied along the last axis."""
    src = np.asarray(src, np.float32)
    src_len = src.shape[-1]
    out = np.empty(src.shape[:-1] + (dst_len,), np.float32)
    if dst_len > src_len:
        if src_len == 1:
            out[...] = src[..., :1]
            return out
        scale = np.float32(src_len - 1) / np.float32(dst_len - 1)
        for di in range(dst_len - 1):
            sf = np.float32(di) * scale
            si = int(sf)
            sf = np.float32(sf - np.float32(si))
            out[..., di] = (np.float32(1) - sf) * src[..., si] + sf * src[..., si + 1]
        out[..., dst_len - 1] = src[..., src_len - 1]
    elif dst_len < src_len:
        scale = np.float32(src_len) / np.float32(dst_len)
        si0_i, si0_f = 0, np.float32(0)
        for di in range(dst_len):
            s1 = np.float32(di + 1) * scale
            si1_i = int(s1)
            si1_f = np.float32(s1 - np.float32(si1_i))
            acc = (np.float32(1) - si0_f) * src[..., si0_i]
            n = np.float32(1) - si0_f
            for si in range(si0_i + 1, si1_i):
                acc = acc + src[..., si]
                n = np.float32(n + np.float32(1))
            if si1_i < src_len:
                acc = acc + si1_f * src[..., si1_i]
                n = np.float32(n + si1_f)
            out[..., di] = acc / n
            si0_i, si0_f = si1_i, si1_f
    else:
        out[...] = src
    return out


def image_scale(src_chw, size):
    """image.scale(src, size) for a (C,H,W) float image and a NUMBER size, bilinear mode (the default)."""
    src = np.asarray(src_chw, np.float32)
    C, H, W = src.shape
    oh, ow = image_scaled_size(H, W, size)
    tmp = _scale_linear_1d(src, ow)                                   # (C, H, ow): width pass
    out = _scale_linear_1d(tmp.transpose(0, 2, 1), oh).transpose(0, 2, 1)   # height pass
    return np.ascontiguousarray(out)


# run_model.lua:66-74: image.load (byte/255) -> image.scale(img, image_size) -> BGR -> *255 -> minus the VGG mean
def preprocess_u8_resize(img_u8_hwc, image_size):
    img = np.asarray(img_u8_hwc, np.uint8)
    x = (img.astype(np.float32) / f32(255)).transpose(2, 0, 1)       # image.load: (3,H,W) RGB in [0,1]
    x = image_scale(x, image_size)
    x = np.ascontiguousarray(x[::-1])                                # index(1, {3,2,1}): BGR
    x = x * f32(255)
    x = x + (f32(-1) * VGG_MEAN_BGR)[:, None, None]
    return x[None].astype(np.float32)


# densecap/box_utils.lua:311-341  (x, y, w, h) -> (x1, y1, x2, y2) with x2 = x + w - 1
def xywh_to_x1y1x2y2(boxes):
    b = _f(boxes)
    out = np.empty_like(b)
    out[..., 0] = b[..., 0]
    out[..., 1] = b[..., 1]
    out[..., 2] = (b[..., 0] + b[..., 2]) + f32(-1)
    out[..., 3] = (b[..., 1] + b[..., 3]) + f32(-1)
    return out


# densecap/modules/InvertBoxTransform.lua:63-97
def invert_box_transform_backward(anchor_boxes, target_boxes, grad_out):
    a = _f(anchor_boxes); t = _f(target_boxes); g = _f(grad_out)
    out = invert_box_transform(a, t)
    ga = np.zeros_like(a); gt = np.zeros_like(t)
    ga[:, 0] = (g[:, 0] / a[:, 2]) * f32(-1)
    ga[:, 1] = (g[:, 1] / a[:, 3]) * f32(-1)
    ga[:, 2] = ((out[:, 0] * g[:, 0] + g[:, 2]) / a[:, 2]) * f32(-1)
    ga[:, 3] = ((out[:, 1] * g[:, 1] + g[:, 3]) / a[:, 3]) * f32(-1)
    gt[:, 0] = g[:, 0] / a[:, 2]
    gt[:, 1] = g[:, 1] / a[:, 3]
    gt[:, 2] = g[:, 2] / t[:, 2]
    gt[:, 3] = g[:, 3] / t[:, 3]
    return ga, gt


# densecap/modules/BoxRegressionCriterion.lua:50-79: w * SmoothL1(transforms, InvertBoxTransform(anchors, targets)) with the
# |t| > 10 mask hack; backward w.r.t. the transforms and (through InvertBoxTransform backward, negated) the anchors.
def box_regression_criterion(anchor_boxes, transforms, target_boxes, w=1.0):
    tt = invert_box_transform(anchor_boxes, target_boxes).copy()
    tr = _f(transforms).copy()
    big = np.abs(tt).max(axis=1) > 10
    tr[big] = 0; tt[big] = 0
    loss, g = smooth_l1(tr, tt)
    loss = f32(w) * loss
    grad_trans = (g * f32(w)).astype(_DT[0])
    din_anchor, _ = invert_box_transform_backward(anchor_boxes, target_boxes, grad_trans)
    return loss, (din_anchor * f32(-1)).astype(_DT[0]), grad_trans


# densecap/LanguageModel.lua:148-164  getTarget: (N,T) ground-truth tokens (0 = padding) -> (N,T+2) targets: a leading 0 for the
# image time step, the tokens, the first padding slot replaced by END (= V+1).  Pinned by test/LanguageModel_test.lua:46-70.
def lm_get_target(gt_sequence, vocab_size):
    gt = np.asarray(gt_sequence, np.int64)
    N, T = gt.shape
    target = np.zeros((N, T + 2), np.int64)
    target[:, 1:T + 1] = gt
    for i in range(N):
        for t in range(1, T + 2):
            if target[i, t] == 0:
                target[i, t] = vocab_size + 1
                break
    return target


def log_softmax(x):
    x = _f(x)
    m = x.max(axis=1, keepdims=True)
    return ((x - m) - np.log(np.exp((x - m).astype(np.float64)).sum(axis=1, keepdims=True)).astype(_DT[0])).astype(_DT[0])


def _topk_desc(v, k):
    """torch.topk(v, k, dim, true): the k largest, descending.  The order among exactly equal values is unspecified in TH;
    the oracle DEFINES it as lower index first (stable)."""
    idx = np.argsort(-v, kind="stable")[:k]
    return v[idx], idx


# densecap/LanguageModel.lua:170-290  LM:beamsearch -- each image vector separately, beams in the minibatch dimension.
# Returns (N,T) int64 tokens (1-based).  SURVEY.md 8f rank 3; the GPU implementation is later work, this is its oracle.
def lm_beamsearch(codes, weights, seq_length, beam_size=20):
    codes = _f(codes)
    N = codes.shape[0]
    Wl = weights["lm.lstm.weight"]; bl = weights["lm.lstm.bias"]
    H = Wl.shape[1] // 4
    V1 = weights["lm.vocab.weight"].shape[0]
    START = END = V1
    T = seq_length
    lookup = _f(weights["lm.lookup.weight"])
    seq = np.zeros((N, T), np.int64)
    for i in range(N):
        enc = relu(linear(codes[i:i + 1], weights["lm.encoder.weight"], weights["lm.encoder.bias"]))
        h = np.zeros((1, H), _DT[0]); c = np.zeros((1, H), _DT[0])
        h, c = lstm_step(enc, h, c, Wl, bl)                                            # image time step
        h, c = lstm_step(lookup[np.array([START]) - 1], h, c, Wl, bl)                  # START token
        scores = linear(h, weights["lm.vocab.weight"], weights["lm.vocab.bias"])
        all_logprobs = log_softmax(scores)[0]
        beam_logprobs, idx = _topk_desc(all_logprobs, beam_size)
        B = idx.shape[0]                                                               # = beam_size (V+1 >= beam_size assumed, as in the reference)
        beams = np.ones((B, T), np.int64)
        beams[:, 0] = idx + 1
        h = np.repeat(h, B, axis=0); c = np.repeat(c, B, axis=0)                        # duplicate the states for all beams (:221-227)
        for t in range(1, T):
            words = beams[:, t - 1]
            h, c = lstm_step(lookup[words - 1], h, c, Wl, bl)
            next_logprobs = log_softmax(linear(h, weights["lm.vocab.weight"], weights["lm.vocab.bias"]))
            # a beam that already holds an END token contributes nothing more (:243-247)
            end_mask = ((beams == END).sum(axis=1) == 0).astype(_DT[0])
            next_logprobs = next_logprobs * end_mask[:, None]
            top_lp = np.empty((B, B), _DT[0]); top_ix = np.empty((B, B), np.int64)
            for b in range(B):
                top_lp[b], top_ix[b] = _topk_desc(next_logprobs[b], B)
            all_next = (top_lp + beam_logprobs[:, None]).reshape(-1)                    # index b*B + k
            beam_logprobs, sel = _topk_desc(all_next, B)
            src = sel // B
            new_beams = beams[src].copy()
            new_beams[:, t] = top_ix.reshape(-1)[sel] + 1
            beams = new_beams
            h = h[src]; c = c[src]
        seq[i] = beams[int(np.argmax(beam_logprobs))]
    return seq
