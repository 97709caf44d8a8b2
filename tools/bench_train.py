"""Times the BASELINE configs[4] training step of the LocalizationLayer (720x600 features 38x45x512, 50 GT boxes,
256 sampled ROIs): forward, backward, and both, with CUDA events.  Prints one JSON line."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densecap_b200.model import LocalizationLayer          # noqa: E402
from densecap_b200.runtime import Context                   # noqa: E402
from densecap_b200.synth import ModelDims, feature_hw       # noqa: E402


def main(steps=20, warmup=3):
    dims = ModelDims()
    rng = np.random.RandomState(5)
    k, F, C5 = dims.num_anchors, dims.rpn_filters, dims.widths[4]
    weights = {"rpn.conv.weight": (rng.randn(F, C5, 3, 3) * 0.01).astype(np.float32), "rpn.conv.bias": np.zeros(F, np.float32),
               "rpn.box.weight": (rng.randn(4 * k, F, 1, 1) * 0.01).astype(np.float32), "rpn.box.bias": np.zeros(4 * k, np.float32),
               "rpn.score.weight": (rng.randn(2 * k, F, 1, 1) * 0.01).astype(np.float32), "rpn.score.bias": np.zeros(2 * k, np.float32)}
    H, W = 600, 720
    fh, fw = feature_hw(H, W)
    feats = torch.from_numpy(np.maximum(rng.randn(1, C5, fh, fw), 0).astype(np.float32)).cuda()
    gt = np.stack([np.array([rng.uniform(100, 620), rng.uniform(100, 500), rng.uniform(40, 200), rng.uniform(40, 200)], np.float32)
                   for _ in range(50)])[None]
    gt_t = torch.from_numpy(gt).cuda()
    labels = torch.zeros((1, 50, 15), dtype=torch.int64)
    ctx = Context(dims, max_image_hw=(H, W), max_proposals=256)
    ctx.load_weights(weights)
    ll = LocalizationLayer(ctx, {"sampler_batch_size": 256, "sampler_low_thresh": 0.3, "box_reg_decay": 5e-5})
    ll.training()
    g1 = torch.randn(256, C5, 7, 7, device="cuda") * 0.01
    g2 = torch.randn(256, 4, device="cuda") * 0.01
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tb = [], []
    for it in range(warmup + steps):
        ll.setImageSize(H, W)
        ll.setGroundTruth(gt_t, labels)
        ev[0].record()
        ll.forward(feats)
        ev[1].record()
        ll.backward(feats, [g1, g2])
        ev[2].record()
        torch.cuda.synchronize()
        if it >= warmup:
            tf.append(ev[0].elapsed_time(ev[1])); tb.append(ev[1].elapsed_time(ev[2]))
    print(json.dumps({"workload": "LocalizationLayer training step, 38x45x512 features, 50 GT boxes, 256 sampled ROIs (BASELINE configs[4])",
                      "forward_ms": float(np.median(tf)), "backward_ms": float(np.median(tb)), "step_ms": float(np.median(tf) + np.median(tb)),
                      "steps": steps, "num_pos": ll.stats["vars"]["num_pos"], "num_neg": ll.stats["vars"]["num_neg"],
                      "note": "forward includes the host read-back of the 5 losses and the sample counts (stats), like the reference's stats table"}))
    ctx.close()


if __name__ == "__main__":
    main()
