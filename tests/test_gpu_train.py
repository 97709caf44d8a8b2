"""GPU parity of the training-time forward of nn.LocalizationLayer (LocalizationLayer.lua:383-526, BASELINE configs[4])
against oracle.localization_layer_forward_train: sampler indices bit-exact, boxes / ROI features / losses within the
north-star fp32 tolerance (1e-4 relative).  Sampling is the deterministic injection both sides implement (first P
positives / first M negatives in index order, the reference's debug_*_sample_idx hook)."""
import numpy as np
import pytest
import torch

from oracle import densecap_oracle as O
from densecap_b200.synth import ModelDims, make_weights, feature_hw

pytestmark = pytest.mark.gpu

SMALL = dict(widths=(64, 64, 64, 64, 64), rpn_filters=64, fc_dim=128, rnn_size=64, encoding_size=64, vocab_size=200, seq_length=5)
IMG_H, IMG_W = 96, 128


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30)


def make_case(dims, weights, img_h, img_w, num_gt, seed, tiny_gt=False):
    """Random non-negative features + GT boxes: half of them copies of in-bounds RPN boxes nudged by a pixel (so real
    positives with IoU > 0.7 exist), the rest random boxes inside the image."""
    rng = np.random.RandomState(seed)
    fh, fw = feature_hw(img_h, img_w)
    feats = np.maximum(rng.randn(1, dims.widths[4], fh, fw), 0).astype(np.float32)
    boxes, _, _, _, _ = O.rpn_forward(feats, weights)
    c = O.xcycwh_to_x1y1x2y2(boxes)[0]
    inb = np.nonzero((c[:, 0] >= 1) & (c[:, 1] >= 1) & (c[:, 2] <= img_w) & (c[:, 3] <= img_h))[0]
    gt = []
    pick = rng.choice(inb, size=min(num_gt // 2, inb.size), replace=False) if inb.size else []
    for r in pick:
        b = boxes[0, r].copy()
        b[:2] += rng.uniform(-1, 1, 2).astype(np.float32)
        gt.append(b)
    while len(gt) < num_gt:
        w = rng.uniform(10, img_w / 2); h = rng.uniform(10, img_h / 2)
        gt.append(np.array([rng.uniform(w / 2, img_w - w / 2), rng.uniform(h / 2, img_h - h / 2), w, h], np.float32))
    if tiny_gt:
        # A degenerate GT box first: zero IoU with everything, so its forced positive is row 1 (first max of an all-zero
        # column); the other GT boxes are kept in the lower-right corner, away from row 1 (the top-left anchor), so row 1's
        # target is the degenerate box and log(w/wa) < -10 fires the |t| > 10 mask hack.
        gt = [np.array([img_w - 2, img_h - 2, 1e-3, 1e-3], np.float32)]
        while len(gt) < num_gt:
            gt.append(np.array([rng.uniform(0.8, 0.9) * img_w, rng.uniform(0.8, 0.9) * img_h, 0.15 * img_w, 0.15 * img_h], np.float32))
    return feats, np.stack(gt)[None].astype(np.float32)


def run_case(ll, feats, gt, img_h, img_w, L=5):
    labels = torch.arange(gt.shape[1] * L, dtype=torch.int64).reshape(1, gt.shape[1], L)
    ll.training()
    ll.setImageSize(img_h, img_w)
    ll.setGroundTruth(dev(gt), labels)
    out = ll.forward(dev(feats))
    torch.cuda.synchronize()
    return out, labels


def check_case(ll, out, labels, ref, feat_tol=1e-4):
    P, M = ref["num_pos"], ref["num_neg"]
    assert (ll.stats["vars"]["num_pos"], ll.stats["vars"]["num_neg"]) == (P, M)
    assert ll.pos_input_idx.cpu().tolist() == ref["pos_input_idx"].tolist()
    assert ll.neg_input_idx.cpu().tolist() == ref["neg_input_idx"].tolist()
    assert ll.pos_target_idx.cpu().tolist() == ref["pos_target_idx"].tolist()
    assert ll.stats["vars"]["total_pos"] == int(ref["pos_mask"].sum())
    assert rel(out[1].cpu().numpy(), ref["roi_boxes"]) < 1e-4
    assert rel(out[0].cpu().numpy(), ref["roi_features"]) < feat_tol
    assert np.array_equal(out[2].cpu().numpy(), ref["pos_target_boxes"])          # pure gather of the GT boxes
    assert out[3].cpu().tolist() == labels[0][ref["pos_target_idx"] - 1].tolist()
    for name, want in ref["losses"].items():
        got = ll.stats["losses"][name]
        assert abs(got - float(want)) <= 1e-4 * max(abs(float(want)), 1e-3), (name, got, float(want))


@pytest.fixture(scope="module")
def small():
    dims = ModelDims(**SMALL)
    return dims, make_weights(dims, seed=3, box_variant="B")


@pytest.fixture(scope="module")
def layer(small):
    from densecap_b200.model import LocalizationLayer
    from densecap_b200.runtime import Context
    dims, weights = small
    ctx = Context(dims, max_image_hw=(IMG_H, IMG_W), max_proposals=64)
    ctx.load_weights({k: v for k, v in weights.items() if k.startswith("rpn.")})
    ll = LocalizationLayer(ctx, {"sampler_batch_size": 32, "mid_box_reg_weight": 0.05, "mid_objectness_weight": 0.1, "box_reg_decay": 5e-5,
                                 "sampler_mode": "deterministic"})
    yield ll
    ctx.close()


@pytest.mark.parametrize("seed,num_gt,tiny", [(0, 6, False), (1, 3, False), (2, 4, True)])
def test_train_forward_matches_oracle(layer, small, seed, num_gt, tiny):
    dims, weights = small
    feats, gt = make_case(dims, weights, IMG_H, IMG_W, num_gt, seed, tiny_gt=tiny)
    ref = O.localization_layer_forward_train(feats, weights, gt, IMG_H, IMG_W, mid_box_reg_weight=0.05, mid_objectness_weight=0.1,
                                             box_reg_decay=5e-5, batch_size=32)
    if tiny:
        assert ref["masked"] >= 1, "test case must exercise the mask hack"
    out, labels = run_case(layer, feats, gt, IMG_H, IMG_W)
    check_case(layer, out, labels, ref)


def test_train_forward_neg_fallback_and_no_bounds(layer, small):
    """low_thresh = 0 leaves no negatives -> neg_mask = not pos_mask (BoxSampler.lua:111-128); bounds switched off."""
    dims, weights = small
    feats, gt = make_case(dims, weights, IMG_H, IMG_W, 5, 7)
    layer.opt.update({"sampler_low_thresh": 0.0, "train_remove_outbounds_boxes": 0})
    layer._push_train_args()
    try:
        boxes, anc, trans, scores, _ = O.rpn_forward(feats, weights)
        pi, pt, ni, pos_mask, neg_mask = O.box_sampler(boxes, gt, 0.0, 0.7, 32, None)
        out, labels = run_case(layer, feats, gt, IMG_H, IMG_W)
        assert layer.pos_input_idx.cpu().tolist() == pi.tolist()
        assert layer.neg_input_idx.cpu().tolist() == ni.tolist()
        assert layer.pos_target_idx.cpu().tolist() == pt.tolist()
        assert layer.stats["vars"]["total_neg"] == int(neg_mask.sum())
    finally:
        layer.opt.update({"sampler_low_thresh": 0.5, "train_remove_outbounds_boxes": 1})
        layer._push_train_args()


def test_train_forward_protocol(layer, small):
    from densecap_b200._lib import DenseCapError
    dims, weights = small
    feats, gt = make_case(dims, weights, IMG_H, IMG_W, 4, 3)
    out, labels = run_case(layer, feats, gt, IMG_H, IMG_W)
    layer.setImageSize(IMG_H, IMG_W)
    with pytest.raises(DenseCapError, match="Must call setGroundTruth before training-time forward pass"):
        layer.forward(dev(feats))                                     # LocalizationLayer.lua:385-387
    layer.setGroundTruth(dev(gt), labels)
    layer.forward(dev(feats))
    layer.setGroundTruth(dev(gt), labels)
    with pytest.raises(DenseCapError, match="Must call setImageSize before each forward pass"):
        layer.forward(dev(feats))                                     # :391-393
    layer.train = False


def test_train_forward_full_size():
    """BASELINE configs[4] shape: 720x600 image (45x38 features, 20520 boxes), 50 GT boxes, sampler batch 256, against the oracle
    (the RPN + sampler + 256-box ROI pooling is small enough for the CPU), plus structural properties."""
    from densecap_b200.model import LocalizationLayer
    from densecap_b200.runtime import Context
    dims = ModelDims()
    rng = np.random.RandomState(5)
    k, F, C5 = dims.num_anchors, dims.rpn_filters, dims.widths[4]
    weights = {"rpn.conv.weight": (rng.randn(F, C5, 3, 3) * 0.01).astype(np.float32), "rpn.conv.bias": np.zeros(F, np.float32),
               "rpn.box.weight": (rng.randn(4 * k, F, 1, 1) * 0.01).astype(np.float32), "rpn.box.bias": np.zeros(4 * k, np.float32),
               "rpn.score.weight": (rng.randn(2 * k, F, 1, 1) * 0.01).astype(np.float32), "rpn.score.bias": np.zeros(2 * k, np.float32)}
    H, W = 600, 720
    feats, gt = make_case(dims, weights, H, W, 50, 11)
    ref = O.localization_layer_forward_train(feats, weights, gt, H, W, mid_box_reg_weight=0.05, mid_objectness_weight=0.1,
                                             box_reg_decay=5e-5, sampler_low=0.3, batch_size=256)
    ctx = Context(dims, max_image_hw=(H, W), max_proposals=256)
    try:
        ctx.load_weights(weights)
        ll = LocalizationLayer(ctx, {"sampler_batch_size": 256, "sampler_low_thresh": 0.3, "box_reg_decay": 5e-5, "sampler_mode": "deterministic"})
        out, labels = run_case(ll, feats, gt, H, W)
        P, M = ll.stats["vars"]["num_pos"], ll.stats["vars"]["num_neg"]
        assert P + M == 256 and 1 <= P <= 128
        pi = ll.pos_input_idx.cpu().numpy(); ni = ll.neg_input_idx.cpu().numpy()
        assert np.all(np.diff(pi) > 0) and np.all(np.diff(ni) > 0) and not set(pi.tolist()) & set(ni.tolist())
        assert out[0].shape == (256, C5, 7, 7) and torch.isfinite(out[0]).all()
        check_case(ll, out, labels, ref)
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------------------------
# backward (LocalizationLayer.lua:529-605 + accGradParameters of the RPN)
# ---------------------------------------------------------------------------------------------------------
def run_backward_case(ll, dims, weights, feats, gt, img_h, img_w, S, kw, seed):
    rng = np.random.RandomState(seed)
    C5 = dims.widths[4]
    g1 = (rng.randn(S, C5, 7, 7) * 0.05).astype(np.float32)
    g2 = (rng.randn(S, 4) * 0.05).astype(np.float32)
    fwd = O.localization_layer_forward_train(feats, weights, gt, img_h, img_w, batch_size=S, **kw)
    ref = O.localization_layer_backward_train(feats, weights, gt, img_h, img_w, fwd, g1, g2, mid_box_reg_weight=kw["mid_box_reg_weight"],
                                              mid_objectness_weight=kw["mid_objectness_weight"], box_reg_decay=kw["box_reg_decay"])
    out, labels = run_case(ll, feats, gt, img_h, img_w)
    check_case(ll, out, labels, fwd)
    ll.zeroGradParameters()
    gin = ll.backward(dev(feats), [dev(g1), dev(g2)])
    torch.cuda.synchronize()
    return gin, ll.gradParameters(), ref


def check_backward(gin, gparams, ref, tol=1e-4):
    assert rel(gin.cpu().numpy(), ref["grad_feats"]) < tol
    for name, want in ref["grad_params"].items():
        got = gparams[name].cpu().numpy()
        assert got.shape == want.shape
        assert rel(got, want) < tol, name


@pytest.mark.parametrize("seed,num_gt", [(0, 6), (4, 3)])
def test_train_backward_matches_oracle(layer, small, seed, num_gt):
    dims, weights = small
    feats, gt = make_case(dims, weights, IMG_H, IMG_W, num_gt, seed)
    kw = dict(mid_box_reg_weight=0.05, mid_objectness_weight=0.1, box_reg_decay=5e-5)
    gin, gparams, ref = run_backward_case(layer, dims, weights, feats, gt, IMG_H, IMG_W, 32, kw, seed)
    check_backward(gin, gparams, ref)
    layer.train = False


def test_train_backward_protocol_and_accumulation(layer, small):
    from densecap_b200._lib import DenseCapError
    dims, weights = small
    feats, gt = make_case(dims, weights, IMG_H, IMG_W, 5, 9)
    kw = dict(mid_box_reg_weight=0.05, mid_objectness_weight=0.1, box_reg_decay=5e-5)
    gin, gparams, ref = run_backward_case(layer, dims, weights, feats, gt, IMG_H, IMG_W, 32, kw, 9)
    g1 = dev(np.zeros((32, dims.widths[4], 7, 7), np.float32)); g2 = dev(np.zeros((32, 4), np.float32))
    # a second backward without re-arming setGroundTruth / setImageSize fails like the reference asserts (:531-537)
    with pytest.raises(DenseCapError, match="Must call setGroundTruth before each forward pass"):
        layer.backward(dev(feats), [g1, g2])
    # gradients accumulate across backward calls until zeroGradParameters (accGradParameters semantics)
    out, labels = run_case(layer, feats, gt, IMG_H, IMG_W)
    rng = np.random.RandomState(9)
    g1 = (rng.randn(32, dims.widths[4], 7, 7) * 0.05).astype(np.float32); g2 = (rng.randn(32, 4) * 0.05).astype(np.float32)
    layer.backward(dev(feats), [dev(g1), dev(g2)])
    twice = layer.gradParameters()
    for name, want in ref["grad_params"].items():
        assert rel(twice[name].cpu().numpy(), 2 * want) < 1e-4, name
    layer.train = False


def test_train_backward_full_size():
    """BASELINE configs[4]: 720x600, 50 GT boxes, 256 sampled ROIs, fwd + bwd against the oracle."""
    from densecap_b200.model import LocalizationLayer
    from densecap_b200.runtime import Context
    dims = ModelDims()
    rng = np.random.RandomState(5)
    k, F, C5 = dims.num_anchors, dims.rpn_filters, dims.widths[4]
    weights = {"rpn.conv.weight": (rng.randn(F, C5, 3, 3) * 0.01).astype(np.float32), "rpn.conv.bias": np.zeros(F, np.float32),
               "rpn.box.weight": (rng.randn(4 * k, F, 1, 1) * 0.01).astype(np.float32), "rpn.box.bias": np.zeros(4 * k, np.float32),
               "rpn.score.weight": (rng.randn(2 * k, F, 1, 1) * 0.01).astype(np.float32), "rpn.score.bias": np.zeros(2 * k, np.float32)}
    H, W = 600, 720
    feats, gt = make_case(dims, weights, H, W, 50, 11)
    kw = dict(mid_box_reg_weight=0.05, mid_objectness_weight=0.1, box_reg_decay=5e-5, sampler_low=0.3)
    ctx = Context(dims, max_image_hw=(H, W), max_proposals=256)
    try:
        ctx.load_weights(weights)
        ll = LocalizationLayer(ctx, {"sampler_batch_size": 256, "sampler_low_thresh": 0.3, "box_reg_decay": 5e-5, "sampler_mode": "deterministic"})
        gin, gparams, ref = run_backward_case(ll, dims, weights, feats, gt, H, W, 256, kw, 3)
        check_backward(gin, gparams, ref)
    finally:
        ctx.close()


def test_random_sampler_draws_uniform_minibatches_the_oracle_accepts(small):
    """nn.BoxSampler draws its minibatch with torch.multinomial (BoxSampler.lua:138-152).  The library's default does the same with a
    seeded generator: every draw is a set of DISTINCT positives / negatives from the oracle's candidate masks, different seeds and
    consecutive calls give different minibatches, the same seed replays, and -- injected into the oracle through the reference's
    debug_*_sample_idx hook -- the whole forward matches."""
    from densecap_b200.model import LocalizationLayer
    from densecap_b200.runtime import Context
    dims, weights = small
    feats, gt = make_case(dims, weights, IMG_H, IMG_W, 6, 0)
    ctx = Context(dims, max_image_hw=(IMG_H, IMG_W), max_proposals=64)
    ctx.load_weights({k: v for k, v in weights.items() if k.startswith("rpn.")})
    try:
        ll = LocalizationLayer(ctx, {"sampler_batch_size": 32, "mid_box_reg_weight": 0.05, "mid_objectness_weight": 0.1, "box_reg_decay": 5e-5,
                                     "sampler_mode": "random", "sampler_seed": 7})
        base = O.localization_layer_forward_train(feats, weights, gt, IMG_H, IMG_W, mid_box_reg_weight=0.05, mid_objectness_weight=0.1,
                                                 box_reg_decay=5e-5, batch_size=32)
        pos_nz = np.nonzero(base["pos_mask"])[0]; neg_nz = np.nonzero(base["neg_mask"])[0]
        draws = []
        for call in range(3):
            out, labels = run_case(ll, feats, gt, IMG_H, IMG_W)
            pi = ll.pos_input_idx.cpu().numpy() - 1; ni = ll.neg_input_idx.cpu().numpy() - 1
            assert len(pi) == base["num_pos"] and len(ni) == base["num_neg"]
            assert set(pi.tolist()) <= set(pos_nz.tolist()) and len(set(pi.tolist())) == len(pi)
            assert set(ni.tolist()) <= set(neg_nz.tolist())
            if neg_nz.size >= len(ni):
                assert len(set(ni.tolist())) == len(ni), "negatives are drawn without replacement when there are enough"
            # the same draw injected into the oracle -> identical forward
            ref = O.localization_layer_forward_train(feats, weights, gt, IMG_H, IMG_W, mid_box_reg_weight=0.05, mid_objectness_weight=0.1,
                                                     box_reg_decay=5e-5, batch_size=32, pos_sample_idx=np.searchsorted(pos_nz, pi),
                                                     neg_sample_idx=np.searchsorted(neg_nz, ni))
            check_case(ll, out, labels, ref)
            draws.append((tuple(pi.tolist()), tuple(ni.tolist())))
        assert len(set(d[1] for d in draws)) > 1, "consecutive calls must draw different minibatches"
        ctx.set_sampler("random", 7)                      # same seed: the sequence replays
        run_case(ll, feats, gt, IMG_H, IMG_W)
        assert (tuple((ll.pos_input_idx.cpu().numpy() - 1).tolist()), tuple((ll.neg_input_idx.cpu().numpy() - 1).tolist())) == draws[0]
        ctx.set_sampler("random", 8)
        run_case(ll, feats, gt, IMG_H, IMG_W)
        assert tuple((ll.neg_input_idx.cpu().numpy() - 1).tolist()) != draws[0][1]
    finally:
        ctx.close()
