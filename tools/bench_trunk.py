"""Device times of the dense stages (conv trunk, fc6, fc7, caption decode) of the full-size model at batch 1 and 4, for A/B runs of
variant builds: DENSECAP_B200_LIB=<variant .so> python tools/bench_trunk.py [groups64 groups128 groups256]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densecap_b200.model import DenseCapModel
from densecap_b200.synth import ModelDims, make_weights, make_image

H, W = 600, 720
dims = ModelDims()
w = make_weights(dims, seed=0, box_variant="C")
img = make_image(H, W, seed=0)
m = DenseCapModel({"dims": dims}, weights=w, max_batch=4)
m.setTestArgs({"rpn_nms_thresh": 0.7, "final_nms_thresh": 0.3, "num_proposals": 1000, "seq_length": 15})
ctx, lib = m.ctx, m.ctx._lib
if len(sys.argv) >= 4:
    ctx.check(lib.dc_set_accumulation_groups(ctx.handle, int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])))
dimg = torch.from_numpy(img).cuda()
out = {"lib": os.environ.get("DENSECAP_B200_LIB", "default"), "groups": sys.argv[1:4]}
for batch in (1, 4):
    x = dimg.repeat(batch, 1, 1, 1).contiguous() if batch > 1 else dimg
    for it in range(3):
        m.forward_device(x)
    ctx.enable_timing(True)
    acc = {}
    for it in range(7):
        m.forward_device(x)
        torch.cuda.synchronize()
        for n_, ms_ in ctx.get_times():
            acc.setdefault(n_, []).append(ms_)
    ctx.enable_timing(False)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for it in range(10):
        m.forward_device(x)
    ev1.record()
    torch.cuda.synchronize()
    out["b%d" % batch] = {k: round(float(np.median(v)), 4) for k, v in acc.items() if k in ("conv_trunk", "recog:fc6", "recog:fc7", "lm:sample", "nms+roi_pooling")}
    out["b%d" % batch]["step_ms"] = round(ev0.elapsed_time(ev1) / 10, 4)
    if os.environ.get("DC_LAYERS"):
        ctx.check(lib.dc_enable_timing(ctx.handle, 2))
        acc = {}
        for it in range(5):
            m.forward_device(x)
            torch.cuda.synchronize()
            for n_, ms_ in ctx.get_times():
                acc.setdefault(n_, []).append(ms_)
        ctx.enable_timing(False)
        out["b%d_layers_us" % batch] = {k: round(1000 * float(np.median(v)), 1) for k, v in acc.items() if k.startswith("conv")}
print(json.dumps(out), flush=True)
