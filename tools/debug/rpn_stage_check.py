"""Which stage of the RPN post-processing deviates from the oracle at full size?  Leaf ops on the GPU, one by one, on the oracle's data."""
import ctypes as C
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import densecap_oracle as O
from densecap_b200.synth import ModelDims, make_weights, make_image, feature_hw
from densecap_b200.runtime import Context

dims = ModelDims()
ctx = Context(dims)
lib = ctx._lib
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
p = lambda t: C.c_void_p(t.data_ptr())
H, W = 600, 720
fh, fw = feature_hw(H, W)
for variant in ("C", "A"):
    w = make_weights(dims, seed=0, box_variant=variant)
    img = make_image(H, W, seed=0)
    feats = O.vgg_trunk(img, w)
    boxes, anc, trans, scores, hid = O.rpn_forward(feats, w)
    n = boxes.shape[1]
    # ApplyBoxTransform
    g = torch.empty((n, 4), device="cuda")
    lib.dc_apply_box_transform(p(dev(anc[0])), p(dev(trans[0])), p(g), n, None)
    print(variant, "apply_box_transform equal:", np.array_equal(g.cpu().numpy(), boxes[0]))
    cb, valid = O.clip_boxes(boxes, 1, 1, W, H, "xcycwh")
    gc = torch.empty((n, 4), device="cuda"); gv = torch.empty((n,), dtype=torch.uint8, device="cuda")
    lib.dc_clip_boxes(p(dev(boxes[0])), p(gc), p(gv), n, 1.0, 1.0, float(W), float(H), None)
    print(variant, "clip equal:", np.array_equal(gc.cpu().numpy(), cb[0]), "valid equal:", np.array_equal(gv.cpu().numpy().astype(bool), valid.astype(bool)), "n_valid", int(valid.sum()))
    m = valid.astype(bool)
    x1 = O.xcycwh_to_x1y1x2y2(cb[:, m])[0]
    e = O.exp_f32(scores[:, m]); pos, neg = e[0, :, 0], e[0, :, 1]
    prob = (np.float32(1) / (pos + neg)) * pos
    gp = torch.empty((int(m.sum()),), device="cuda")
    lib.dc_objectness_prob(p(dev(scores[0, m])), p(gp), int(m.sum()), None)
    gpn = gp.cpu().numpy()
    print(variant, "prob equal:", np.array_equal(gpn, prob), "mismatches", int((gpn != prob).sum()), "max abs", float(np.abs(gpn - prob).max()))
    bs = np.concatenate([x1, prob[:, None]], 1).astype(np.float32)
    ref = O.nms(bs, 0.7, 1000)
    idx = torch.zeros((1000,), dtype=torch.int64, device="cuda"); cnt = torch.zeros((1,), dtype=torch.int32, device="cuda")
    ctx.check(lib.dc_nms(ctx.handle, p(dev(bs)), bs.shape[0], 0.7, 1000, p(idx), p(cnt), None))
    k = int(cnt.item()); gi = idx[:k].cpu().numpy()
    same = np.array_equal(gi, ref)
    first = int(np.argmax(gi[:min(k, len(ref))] != ref[:min(k, len(ref))])) if not same else -1
    print(variant, "dc_nms on oracle boxes/probs: count", k, "ref", len(ref), "identical", same, "first diff at", first)
    if not same:
        print("   gpu", gi[max(0, first - 2):first + 4], "ref", ref[max(0, first - 2):first + 4])
        order = np.argsort(prob, kind="stable")[::-1]
        rank = np.empty_like(order); rank[order] = np.arange(len(order))
        print("   sorted position of the differing picks: gpu", rank[gi[first] - 1], "ref", rank[ref[first] - 1], "deepest ref pick", rank[ref - 1].max())
    # with the GPU's own prob (in case prob differs)
    bs2 = bs.copy(); bs2[:, 4] = gpn
    ref2 = O.nms(bs2, 0.7, 1000)
    print(variant, "oracle nms on GPU probs identical to GPU nms:", np.array_equal(gi, ref2))
ctx.close()
