"""Evaluator core (eval/eval_utils.lua:134-312, box_utils.merge_boxes) -- host logic, checked on hand-built cases and against the
reference's own smoke test (test/evaluation_test.lua: evaluatorTest must run and return the four result fields)."""
import numpy as np

from densecap_b200.evaluation import DenseCaptioningEvaluator, iou_matrix, merge_boxes, xcycwh_to_x1y1x2y2


def test_iou_matrix_and_merge_groups():
    boxes = np.array([[1, 1, 10, 10], [1, 1, 10, 11], [50, 50, 60, 60], [2, 1, 10, 10], [100, 100, 101, 101]], np.float32)
    D = iou_matrix(boxes)
    assert D.shape == (5, 5) and np.allclose(np.diag(D), 1) and D[0, 2] == 0 and abs(D[0, 1] - 100.0 / 110.0) < 1e-12
    groups = merge_boxes(boxes, 0.7)
    assert [g.tolist() for g in groups] == [[0, 1, 3], [2], [4]]          # the box with most neighbours first, then the singletons in order


def test_perfect_detections_give_map_one_and_misses_lower_it():
    gt = np.array([[20, 20, 10, 10], [60, 60, 20, 20], [100, 30, 16, 8]], np.float32)          # xcycwh, no overlaps
    texts = ["a red car", "the tall tree", "some sky"]
    ev = DenseCaptioningEvaluator()
    ev.addResult(np.array([[0.9, 0.1], [0.8, 0.2], [0.7, 0.3]]), gt, texts, gt, texts)
    res = ev.evaluate()
    assert ev.numAdded() == 1 and abs(res["map"] - 1.0) < 1e-12 and abs(res["detmap"] - 1.0) < 1e-12
    # second image: one correct box with the wrong caption, one box far away, one duplicate of an already matched box
    det = np.array([[20, 20, 10, 10], [200, 200, 10, 10], [20, 20, 10, 10]], np.float32)
    ev.addResult(np.array([0.95, 0.5, 0.4]), det, ["a blue car", "nothing", "a red car"], gt, texts)
    res2 = ev.evaluate()
    assert res2["detmap"] < 1.0 and res2["map"] < res2["detmap"]                                  # language threshold removes the mis-captioned hit
    assert set(res2) == {"map", "ap_breakdown", "detmap", "det_breakdown"} and len(res2["ap_breakdown"]) == 30 and len(res2["det_breakdown"]) == 5


def test_reference_smoke_case_runs():
    """test/evaluation_test.lua evaluatorTest: random boxes and captions, two images, must produce finite results."""
    rng = np.random.RandomState(0)
    ev = DenseCaptioningEvaluator()
    for B, M in ((10, 5), (7, 4)):
        ev.addResult(rng.randn(B, 2), rng.rand(B, 4).astype(np.float32), ["pred %d" % i for i in range(B)],
                     rng.rand(M, 4).astype(np.float32), ["gt %d" % i for i in range(M)])
    res = ev.evaluate()
    assert 0.0 <= res["map"] <= 1.0 and 0.0 <= res["detmap"] <= 1.0 and np.isfinite(res["map"])


def test_box_format_conversion_matches_oracle():
    from oracle import densecap_oracle as O
    b = np.random.RandomState(1).rand(20, 4).astype(np.float32) * 100
    assert np.array_equal(xcycwh_to_x1y1x2y2(b), O.xcycwh_to_x1y1x2y2(b))


def test_unmatched_detection_quirk_consumes_last_ground_truth():
    """eval_utils.lua:169-205 with jmax = -1: `used[-1]` is the last merged GT box in Torch7.  A confident detection that overlaps
    nothing therefore blocks the later, correct detection of the LAST ground-truth box (reference behaviour, kept by default)."""
    gt = np.array([[20, 20, 10, 10], [60, 60, 20, 20]], np.float32)
    texts = ["first", "second"]
    det = np.array([[300, 300, 10, 10], [60, 60, 20, 20], [20, 20, 10, 10]], np.float32)
    for quirks, expect_ok in ((True, [1, 0, 1]), (False, [0, 1, 1])):
        ev = DenseCaptioningEvaluator(reference_quirks=quirks)
        ev.addResult(np.array([0.9, 0.8, 0.7]), det, ["junk", "second", "first"], gt, texts)
        assert [r["ok"] for r in ev.records] == expect_ok
    faithful = DenseCaptioningEvaluator(reference_quirks=True); intended = DenseCaptioningEvaluator(reference_quirks=False)
    for ev in (faithful, intended):
        ev.addResult(np.array([0.9, 0.8, 0.7]), det, ["junk", "second", "first"], gt, texts)
    assert faithful.evaluate()["detmap"] < intended.evaluate()["detmap"]


def test_ap_sweep_reproduces_the_lua_float_loop():
    """eval_utils.lua:285-295: `for t=0,1,0.01` accumulates the step in double precision, so the sweep has 100 thresholds
    (the last is ~0.99, 1.0000000000000007 > 1 ends the loop) and threshold #50 is 0.5000000000000002 > 0.5.
    Hand-computed case: 2 ground-truth boxes; detections in score order = hit, miss, hit -> rec = [.5, .5, 1], prec = [1, .5, 2/3].
    Thresholds 0..49 (<= 0.5) see max precision 1, thresholds 50..99 see 2/3: AP = (50 + 50 * 2/3) / 100 = 0.8333...
    (a 101-point k*0.01 sweep would give (51 + 50 * 2/3) / 101 = 0.83498)."""
    gt = np.array([[20, 20, 10, 10], [60, 60, 20, 20]], np.float32)
    det = np.array([[20, 20, 10, 10], [300, 300, 10, 10], [60, 60, 20, 20]], np.float32)
    ev = DenseCaptioningEvaluator(reference_quirks=False)
    ev.addResult(np.array([0.9, 0.8, 0.7]), det, ["a", "b", "c"], gt, ["a", "c"])
    res = ev.evaluate()
    want = (50 + 50 * (2.0 / 3.0)) / 100
    assert all(abs(v - want) < 1e-12 for v in res["det_breakdown"].values()), res["det_breakdown"]
    assert abs(res["detmap"] - want) < 1e-12
