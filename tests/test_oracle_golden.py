"""Pins the CPU oracle against the known-answer vectors held by the reference's own tests
(tests/golden/reference_known_answers.json, transcribed by make_reference_known_answers.py).
CPU only."""
import json
import os

import numpy as np

from oracle import densecap_oracle as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))


def test_nms_known_answers():
    for c in G["nms"]:
        bs = np.concatenate([np.array(c["boxes"], np.float32), np.array(c["scores"], np.float32)[:, None]], 1)
        idx = O.nms(bs, c["thresh"])
        assert idx.tolist() == c["expected_idx"], c["src"]
        assert idx.dtype == np.int64


def test_nms_max_boxes_and_empty():
    c = G["nms"][0]
    bs = np.concatenate([np.array(c["boxes"], np.float32), np.array(c["scores"], np.float32)[:, None]], 1)
    assert O.nms(bs, 0.7, 3).tolist() == c["expected_idx"][:3]
    assert O.nms(np.zeros((0, 5), np.float32), 0.7).shape == (0,)
    assert O.nms(bs[:1], 0.7).tolist() == [1]


def test_box_to_affine_known_answer():
    c = G["box_to_affine"]
    th = O.box_to_affine(np.array(c["boxes"], np.float32), c["H"], c["W"])
    np.testing.assert_allclose(th, np.array(c["expected"], np.float32), atol=c["tol"], rtol=0)


def test_apply_box_transform_known_answer():
    c = G["apply_box_transform"]
    out = O.apply_box_transform(np.array(c["boxes"], np.float32), np.array(c["trans"], np.float32))
    np.testing.assert_allclose(out, np.array(c["expected"], np.float32), atol=c["tol"], rtol=0)


def _anchors_pipeline(c):
    N, k, H, W = c["N"], c["k"], c["H"], c["W"]
    anchors = np.array(c["anchors"], np.float32)
    trans = np.zeros((N, 4 * k, H, W), np.float32)
    for (n, a, h, w, v) in c["inputs"]:
        trans[n, 4 * a:4 * a + 4, h, w] = v
    anc = O.reshape_box_features(O.make_anchors(N, H, W, c["x0"], c["y0"], c["sx"], c["sy"], anchors), k)
    tr = O.reshape_box_features(trans, k)
    return O.apply_box_transform(anc, tr)          # (N, kHW, 4)


def test_make_boxes_known_answers_and_row_order():
    for c in G["make_boxes"]:
        boxes = _anchors_pipeline(c)
        H, W = c["H"], c["W"]
        for (n, h, w, a, v) in c["expected"]:
            r = (a * H + h) * W + w
            np.testing.assert_allclose(boxes[n, r], np.array(v, np.float32), atol=c["tol"], rtol=0, err_msg=c["src"])


def test_reshape_consistency():
    c = G["reshape_consistency"]
    N, k, D, H, W = c["N"], c["k"], c["D"], c["H"], c["W"]
    trans = np.zeros((N, 4 * k, H, W), np.float32)
    feats = np.zeros((N, D * k, H, W), np.float32)
    trans[c["n"], 4 * c["a"]:4 * c["a"] + 4, c["y"], c["x"]] = c["trans"]
    feats[c["n"], D * c["a"]:D * c["a"] + D, c["y"], c["x"]] = c["feat_fill"]
    anchors = np.array(c["anchors"], np.float32)
    anc = O.reshape_box_features(O.make_anchors(N, H, W, c["x0"], c["y0"], c["sx"], c["sy"], anchors), k)
    boxes = O.apply_box_transform(anc, O.reshape_box_features(trans, k))
    fo = O.reshape_box_features(feats, k)
    r = c["expected_row_1based"] - 1
    assert boxes[c["n"], r].tolist() == c["expected_box"]
    assert fo[c["n"], r].tolist() == [c["feat_fill"]] * D
    # backward is the inverse permutation
    g = np.random.default_rng(0).standard_normal(fo.shape).astype(np.float32)
    gi = O.reshape_box_features_backward(g, k, H, W)
    assert np.array_equal(O.reshape_box_features(gi, k), g)


def test_decode_sequence():
    c = G["decode_sequence"]
    tok = {i + 1: t for i, t in enumerate(c["idx_to_token"])}
    assert O.decode_sequence(np.array(c["seq"]), tok, c["V"] + 1) == c["expected"]


def test_apply_invert_round_trip():
    # test/InvertBoxTransform_test.lua:15-58: Apply(anchors, Invert(anchors, targets)) == targets
    rng = np.random.default_rng(1)
    a = np.concatenate([rng.standard_normal((50, 2)) * 10, rng.uniform(5, 100, (50, 2))], 1).astype(np.float32)
    t = np.concatenate([rng.standard_normal((50, 2)) * 10, rng.uniform(5, 100, (50, 2))], 1).astype(np.float32)
    tr = O.invert_box_transform(a, t)
    back = O.apply_box_transform(a, tr)
    np.testing.assert_allclose(back, t, rtol=1e-5, atol=1e-4)


def test_box_iou_head_semantics_documented_divergence():
    c = G["documented_divergence"]["box_iou"]
    iou = O.box_iou(np.array([[c["box1"]]], np.float32), np.array([[c["box2"]]], np.float32))[0, 0, 0]
    assert abs(iou - c["head_gives"]) < 1e-6
    assert abs(iou - c["stale_test_expects"]) > 1e-2


def test_clip_boxes_head_quirk_width_shrinks_by_one():
    # box_utils.lua:288-291 vs :402-403: round trip shrinks w,h by 1 even for in-bounds boxes
    b = np.array([[[50.0, 40.0, 21.0, 11.0]]], np.float32)
    c, valid = O.clip_boxes(b, 1, 1, 200, 100, "xcycwh")
    assert c[0, 0].tolist() == [50.0, 40.0, 20.0, 10.0] and valid.tolist() == [1]
    # far out-of-bounds box clamps to a 1-px sliver and stays "valid"
    c, valid = O.clip_boxes(np.array([[[-500.0, -500.0, 20.0, 20.0]]], np.float32), 1, 1, 200, 100)
    assert valid.tolist() == [1] and c[0, 0].tolist() == [1.5, 1.5, 1.0, 1.0]


def test_sampler_identity_and_zero_outside():
    # sampling the grid that lands exactly on pixel centres returns the pixels
    H, W, C = 5, 6, 3
    feats = np.random.default_rng(2).standard_normal((H, W, C)).astype(np.float32)
    ys = np.linspace(-1, 1, H).astype(np.float32); xs = np.linspace(-1, 1, W).astype(np.float32)
    grid = np.stack(np.meshgrid(ys, xs, indexing="ij"), -1)[None]
    out = O.bilinear_sample(feats, grid)
    np.testing.assert_allclose(out[0], feats, atol=1e-5)
    far = np.full((1, 2, 2, 2), 3.0, np.float32)
    assert np.all(O.bilinear_sample(feats, far) == 0)


def test_roi_pool_full_image_box_matches_corners():
    C, Hf, Wf = 4, 6, 8
    feats = np.random.default_rng(3).standard_normal((C, Hf, Wf)).astype(np.float32)
    H, W = 96, 128
    box = np.array([[(W + 1) / 2, (H + 1) / 2, W, H]], np.float32)
    out = O.bilinear_roi_pooling(feats, box, H, W, 7, 7)
    assert out.shape == (1, C, 7, 7)
    np.testing.assert_allclose(out[0, :, 0, 0], feats[:, 0, 0], atol=1e-5)
    np.testing.assert_allclose(out[0, :, 6, 6], feats[:, -1, -1], atol=1e-5)


def test_preprocess_u8_restates_run_model():
    """run_model.lua:66-74: image.load float/255, BGR, *255, -mean; values stay within half an ulp of (u8 - mean)."""
    rng = np.random.RandomState(0)
    u8 = rng.randint(0, 256, size=(5, 7, 3), dtype=np.uint8)
    x = O.preprocess_u8(u8)
    assert x.shape == (1, 3, 5, 7) and x.dtype == np.float32
    ideal = u8[..., ::-1].transpose(2, 0, 1).astype(np.float64) - O.VGG_MEAN_BGR.astype(np.float64)[:, None, None]
    assert np.abs(x[0] - ideal).max() < 3e-5
    assert x[0, 0, 2, 3] == np.float32(np.float32(np.float32(u8[2, 3, 2]) / np.float32(255)) * np.float32(255)) + np.float32(-103.939)
