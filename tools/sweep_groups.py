"""Accuracy / speed sweep of the hierarchical-accumulation group size (dc_set_accumulation_groups) on the full-size calibrated model:
conv-trunk and fc6 device times (CUDA events) and the distance of the trunk features / fc7 codes from the fp64 gold run, next to the fp32
oracle's own distance.  Usage (GPU box): python tools/sweep_groups.py"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import densecap_oracle as O
from densecap_b200.model import DenseCapModel
from densecap_b200.synth import ModelDims, make_weights, make_image, feature_hw

H, W = 600, 720
dims = ModelDims()
w = make_weights(dims, seed=0, box_variant="C")
img = make_image(H, W, seed=0)
cfg = {"rpn_nms_thresh": 0.7, "final_nms_thresh": 0.3, "num_proposals": 1000, "seq_length": 15}
ref = O.densecap_forward_test(img, w, cfg, return_intermediates=True)
with O.precision(np.float64):
    gold = O.densecap_forward_test(img, w, cfg, return_intermediates=True)
    gcodes = O.recog_head(ref["ll"]["roi_features"], ref["ll"]["boxes"], w)[0]
rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
floor_feats = rel(ref["feats"], gold["feats"])
floor_codes = rel(O.recog_head(ref["ll"]["roi_features"], ref["ll"]["boxes"], w)[0], gcodes)
m = DenseCapModel({"dims": dims}, weights=w, max_batch=4)
m.setTestArgs(cfg)
ctx, lib = m.ctx, m.ctx._lib
p = lambda t: C.c_void_p(t.data_ptr())
dimg = torch.from_numpy(img).cuda()
dimg4 = dimg.repeat(4, 1, 1, 1).contiguous()
fh, fw = feature_hw(H, W)
feats = torch.empty((1, 512, fh, fw), device="cuda")
roi = torch.from_numpy(ref["ll"]["roi_features"]).cuda(); boxes = torch.from_numpy(ref["ll"]["boxes"]).cuda()
K = roi.shape[0]
codes = torch.empty((K, 4096), device="cuda")
rows = []
for g in [(1, 1, 1), (2, 2, 1), (2, 2, 2), (3, 2, 2), (4, 3, 2), (4, 4, 4), (8, 8, 8), (1000, 1000, 1000)]:
    ctx.check(lib.dc_set_accumulation_groups(ctx.handle, *g))
    ctx.check(lib.dc_vgg_trunk_forward(ctx.handle, p(dimg), 1, H, W, p(feats), None))
    e_feats = rel(feats.cpu().numpy(), gold["feats"])
    ctx.check(lib.dc_recog_forward(ctx.handle, p(roi), p(boxes), K, p(codes), None, None, None, None))
    e_codes = rel(codes.cpu().numpy(), gcodes)
    ph = {}
    for batch, x in ((1, dimg), (4, dimg4)):
        for it in range(3):
            m.forward_device(x)
        ctx.enable_timing(True)
        acc = {}
        for it in range(5):
            m.forward_device(x)
            torch.cuda.synchronize()
            for n_, ms_ in ctx.get_times():
                acc.setdefault(n_, []).append(ms_)
        ctx.enable_timing(False)
        ph[batch] = {k: float(np.median(v)) for k, v in acc.items()}
    rows.append({"groups": g, "trunk_err_vs_fp64": e_feats, "codes_err_vs_fp64": e_codes,
                 "b1": {k: round(ph[1].get(k, 0), 4) for k in ("conv_trunk", "recog:fc6", "recog:fc7", "lm:sample")},
                 "b4": {k: round(ph[4].get(k, 0), 4) for k in ("conv_trunk", "recog:fc6", "recog:fc7", "lm:sample")}})
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps({"oracle_fp32_noise": {"trunk": floor_feats, "codes": floor_codes}}))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"rows": rows, "oracle_fp32_noise": {"trunk": floor_feats, "codes": floor_codes}}, open("gpurun_out/sweep_groups.json", "w"), indent=1)
