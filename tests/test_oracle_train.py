"""Oracle of the training-time LocalizationLayer forward (BASELINE config 5) pinned to the reference's tests that
agree with HEAD: the BoxSamplerHelper gather/scatter protocol (test/BoxSamplerHelper_test.lua:30-104, which injects
the indices through a fake sampler exactly like this restatement does) and the criterion definitions.  CPU only."""
import numpy as np

from oracle import densecap_oracle as O


def test_box_sampler_helper_gather_scatter_protocol():
    # same structure as test/BoxSamplerHelper_test.lua:30-104
    rng = np.random.default_rng(0)
    B1, B2 = 7, 4
    input_data = [rng.standard_normal((1, B1, D)).astype(np.float32) for D in (4, 5, 3)]
    target_data = [rng.standard_normal((1, B2, D)).astype(np.float32) for D in (4, 1)]
    pi = np.array([7, 2, 3]); pt = np.array([1, 2, 1]); ni = np.array([5, 1])
    pos, tgt, neg = O.box_sampler_helper_forward(input_data, target_data, pi, pt, ni)
    for i, D in enumerate((4, 5, 3)):
        assert pos[i].shape == (3, D) and neg[i].shape == (2, D)
        for j in range(3):
            assert np.array_equal(pos[i][j], input_data[i][0, pi[j] - 1])
        for j in range(2):
            assert np.array_equal(neg[i][j], input_data[i][0, ni[j] - 1])
    for i, D in enumerate((4, 1)):
        assert tgt[i].shape == (3, D)
        for j in range(3):
            assert np.array_equal(tgt[i][j], target_data[i][0, pt[j] - 1])
    gp = [rng.standard_normal(p.shape).astype(np.float32) for p in pos]
    gn = [rng.standard_normal(n.shape).astype(np.float32) for n in neg]
    gi = O.box_sampler_helper_backward([d.shape for d in input_data], pi, ni, gp, gn)
    for i in range(3):
        assert gi[i].shape == input_data[i].shape
        for j in range(3):
            assert np.array_equal(gi[i][0, pi[j] - 1], gp[i][j])
        for j in range(2):
            assert np.array_equal(gi[i][0, ni[j] - 1], gn[i][j])
        untouched = [r for r in range(B1) if r + 1 not in set(pi) | set(ni)]
        assert np.all(gi[i][0, untouched] == 0)


def test_box_sampler_masks_and_forced_positives():
    # two GT boxes; inputs: an exact copy of GT 1 (IoU<1 by the HEAD (w-1)/2 convention but > 0.7), a far box (neg),
    # a mediocre overlap (neither), and the best-but-weak match of GT 2 (forced positive, BoxSampler.lua:104-105)
    gt = np.array([[[50, 50, 40, 40], [150, 60, 30, 30]]], np.float32)
    inp = np.array([[[50, 50, 40, 40], [300, 300, 20, 20], [58, 50, 40, 40], [160, 66, 30, 30], [400, 40, 10, 10]]], np.float32)
    pi, pt, ni, pm, nm = O.box_sampler(inp, gt, 0.5, 0.7, batch_size=4)
    iou = O.box_iou(inp, gt)[0]
    assert iou[0, 0] > 0.7 and iou[0, 0] < 1.0            # HEAD quirk: IoU(box, box) < 1
    assert pm.tolist() == [True, False, False, True, False]   # row 3 forced although its IoU < 0.7
    assert iou[3, 1] < 0.7
    assert nm.tolist() == [False, True, False, False, True] and 0.5 <= iou[2, 0] <= 0.7
    assert pi.tolist() == [1, 4] and pt.tolist() == [1, 2]
    assert ni.tolist() == [2, 5]                          # batch 4 -> P = 2, M = 2, first candidates in index order
    # bounds exclusion (BoxSampler.lua:82-98): box 5 sticks out of a 380-wide image
    _, _, ni2, _, nm2 = O.box_sampler(inp, gt, 0.5, 0.7, batch_size=4, bounds={"x_min": 1, "y_min": 1, "x_max": 380, "y_max": 400})
    assert nm2.tolist() == [False, True, False, False, False]
    assert ni2.tolist() == [2, 2]                         # too few negatives -> sampled with replacement


def test_criterions_match_torch_definitions():
    import torch
    rng = np.random.default_rng(1)
    s = rng.standard_normal((9, 2)).astype(np.float32)
    t = rng.integers(1, 3, size=9)
    loss, grad = O.cross_entropy(s, t)
    ts = torch.tensor(s, requires_grad=True)
    tl = torch.nn.functional.cross_entropy(ts, torch.tensor(t - 1))
    tl.backward()
    assert abs(float(loss) - float(tl)) < 1e-6
    np.testing.assert_allclose(grad, ts.grad.numpy(), atol=1e-6)
    x = (rng.standard_normal((6, 4)) * 2).astype(np.float32); y = rng.standard_normal((6, 4)).astype(np.float32)
    loss, grad = O.smooth_l1(x, y)
    tx = torch.tensor(x, requires_grad=True)
    tl = torch.nn.functional.smooth_l1_loss(tx, torch.tensor(y))
    tl.backward()
    assert abs(float(loss) - float(tl)) < 1e-6
    np.testing.assert_allclose(grad, tx.grad.numpy(), atol=1e-6)


def _numeric_gradient(f, x, h=1e-5):
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        old = x[i]
        x[i] = old + h; fp = f(x)
        x[i] = old - h; fm = f(x)
        x[i] = old
        g[i] = (fp - fm) / (2 * h)
    return g


def _relative_error(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(a) + np.abs(b), 1e-12))


def test_box_regression_criterion_gradcheck():
    """test/BoxRegressionCriterion_test.lua: numeric gradients w.r.t. the anchors and the transforms, relative error <= 1e-4
    (run in the oracle's float64 gold mode, like the reference's DoubleTensor test)."""
    rng = np.random.RandomState(0)
    B, w = 10, 2.71
    trans = rng.randn(B, 4)
    anchors = rng.randn(B, 4); anchors[:, 2:] = np.abs(anchors[:, 2:]) + 0.1
    targets = rng.randn(B, 4); targets[:, 2:] = np.abs(targets[:, 2:]) + 0.1
    with O.precision(np.float64):
        loss, g_anchor, g_trans = O.box_regression_criterion(anchors, trans, targets, w)
        f_a = lambda x: float(O.box_regression_criterion(x, trans, targets, w)[0])
        f_t = lambda x: float(O.box_regression_criterion(anchors, x, targets, w)[0])
        assert _relative_error(_numeric_gradient(f_a, anchors.copy()), g_anchor) <= 1e-4
        assert _relative_error(_numeric_gradient(f_t, trans.copy()), g_trans) <= 1e-4


def test_box_conversion_round_trip():
    """test/box_conversion_test.lua: xywh -> x1y1x2y2 -> xywh -> x1y1x2y2 is the identity (1e-6)."""
    rng = np.random.RandomState(1)
    xywh = rng.randn(100, 4).astype(np.float32)
    xywh[:, 2:] = np.abs(xywh[:, 2:])
    a = O.xywh_to_x1y1x2y2(xywh)
    b = O.x1y1x2y2_to_xywh(a)
    c = O.xywh_to_x1y1x2y2(b)
    assert np.abs(a - c).max() <= 1e-6 and np.abs(xywh - b).max() <= 1e-6


def test_lm_get_target_known_answer():
    """test/LanguageModel_test.lua:46-70 (V = 100, T = 5): the reference's expected tensor."""
    gt = np.array([[5, 7, 0, 0, 0], [11, 12, 13, 0, 0], [15, 16, 17, 18, 29], [50, 0, 0, 0, 0]])
    want = np.array([[0, 5, 7, 101, 0, 0, 0], [0, 11, 12, 13, 101, 0, 0], [0, 15, 16, 17, 18, 29, 101], [0, 50, 101, 0, 0, 0, 0]])
    assert np.array_equal(O.lm_get_target(gt, 100), want)


def test_lm_beamsearch_properties():
    """LanguageModel.lua:170-290: with one beam the search is the greedy sample (until an END token: greedy keeps decoding, the beam
    search zeroes the log-probs of a finished beam -- so compare up to the first END); wider beams keep shapes and token range."""
    from densecap_b200.synth import ModelDims, make_weights
    dims = ModelDims(widths=(8, 8, 8, 8, 8), rpn_filters=8, fc_dim=32, rnn_size=16, encoding_size=16, vocab_size=12, seq_length=6)
    w = make_weights(dims, seed=4, box_variant="B")
    codes = np.maximum(np.random.RandomState(0).randn(5, dims.fc_dim), 0).astype(np.float32)
    greedy = O.lm_sample(codes, w, dims.seq_length)
    beam1 = O.lm_beamsearch(codes, w, dims.seq_length, beam_size=1)
    END = dims.vocab_size + 1
    for g, b in zip(greedy, beam1):
        stop = (list(g).index(END) + 1) if END in g else len(g)
        assert list(g[:stop]) == list(b[:stop])
    beam4 = O.lm_beamsearch(codes, w, dims.seq_length, beam_size=4)
    assert beam4.shape == (5, dims.seq_length) and beam4.min() >= 1 and beam4.max() <= END


def test_backward_train_matches_finite_differences():
    """The oracle's LocalizationLayer backward (LocalizationLayer.lua:529-605) against central differences of its own float64 forward,
    on a case without duplicated negatives (with duplicates the reference's indexCopy scatter deliberately keeps only the last copy,
    so its gradient is not the true one)."""
    rng = np.random.RandomState(3)
    C5, F, k, H, W = 8, 8, 12, 6, 8
    img_h, img_w = 96, 128
    w = {"rpn.conv.weight": (rng.randn(F, C5, 3, 3) * 0.05).astype(np.float32), "rpn.conv.bias": (rng.randn(F) * 0.01).astype(np.float32),
         "rpn.box.weight": (rng.randn(4 * k, F, 1, 1) * 0.05).astype(np.float32), "rpn.box.bias": np.zeros(4 * k, np.float32),
         "rpn.score.weight": (rng.randn(2 * k, F, 1, 1) * 0.05).astype(np.float32), "rpn.score.bias": np.zeros(2 * k, np.float32)}
    feats = np.maximum(rng.randn(1, C5, H, W), 0).astype(np.float32)
    gt = np.array([[[30, 30, 40, 50], [90, 60, 50, 40], [64, 48, 30, 30]]], np.float32)
    kw = dict(mid_box_reg_weight=0.5, mid_objectness_weight=1.0, box_reg_decay=1e-3)
    S = 8
    fwd = O.localization_layer_forward_train(feats, w, gt, img_h, img_w, batch_size=S, **kw)
    assert len(set(fwd["neg_input_idx"].tolist())) == fwd["num_neg"], "test case must not repeat negatives"
    g1 = (rng.randn(S, C5, 7, 7)).astype(np.float32); g2 = (rng.randn(S, 4) * 0.1).astype(np.float32)
    bwd = O.localization_layer_backward_train(feats, w, gt, img_h, img_w, fwd, g1, g2, **kw)
    w64 = {n: v.astype(np.float64) for n, v in w.items()}

    def scalar(f):
        with O.precision(np.float64):
            r = O.localization_layer_forward_train(f, w64, gt.astype(np.float64), img_h, img_w, batch_size=S, **kw)
        return float((r["roi_features"] * g1).sum() + (r["roi_boxes"] * g2).sum() + r["losses"]["total_loss"])

    worst = 0.0
    for ix in [(0, 1, 2, 3), (0, 5, 4, 6), (0, 7, 0, 0), (0, 3, 5, 7)]:
        eps = 1e-4
        fp = feats.astype(np.float64).copy(); fp[ix] += eps
        fm = feats.astype(np.float64).copy(); fm[ix] -= eps
        num = (scalar(fp) - scalar(fm)) / (2 * eps)
        ana = float(bwd["grad_feats"][ix])
        worst = max(worst, abs(num - ana) / max(abs(num), abs(ana), 1e-3))
    assert worst < 2e-3, worst
