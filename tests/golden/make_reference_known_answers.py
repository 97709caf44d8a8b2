#!/usr/bin/env python
"""Transcribes the hand-computed known-answer vectors held by the reference's own unit
tests (jcjohnson/densecap @ 7c32170, test/*.lua) into reference_known_answers.json.

These are DATA copied from the reference's tests (inputs + expected outputs), not code;
each case records the file:line it was read from.  Only cases that agree with the HEAD
Lua code are used as pins (SURVEY.md section 4 lists the stale ones, which are kept under
"documented_divergence" with the value HEAD actually produces).

Run:  python tests/golden/make_reference_known_answers.py
"""
import json
import math
import os

L = math.log
cases = {}

# ---- test/nms_test.lua:9-95 -------------------------------------------------------------
_b = [[-12, 3, -7, 9], [-9, 7, -4, 13], [-8, 8, -3, 14], [3.5, 4.5, 8.5, 12.5],
      [-6, -6, -1, -1], [4, 5, 9, 13], [4.5, 5.5, 9.5, 13.5]]
cases["nms"] = [
    {"src": "test/nms_test.lua:9-31", "boxes": _b, "scores": [1, 2, 3, 4, 5, 6, 7], "thresh": 0.7,
     "expected_idx": [7, 5, 4, 3, 2, 1]},
    {"src": "test/nms_test.lua:38-60", "boxes": _b, "scores": [1, 2, 3, 4, 5, 6, 7], "thresh": 0.5,
     "expected_idx": [7, 5, 3, 1]},
    {"src": "test/nms_test.lua:69-91", "boxes": _b, "scores": [2.5, 2, 3, 4, 5, 10, 7], "thresh": 0.7,
     "expected_idx": [6, 5, 3, 1, 2]},
]

# ---- test/BoxToAffine_test.lua:14-44 ----------------------------------------------------
H, W = 20, 30
cases["box_to_affine"] = {
    "src": "test/BoxToAffine_test.lua:14-44", "H": H, "W": W, "tol": 1e-6,
    "boxes": [[(W + 1) / 2, (H + 1) / 2, W, H], [10, 12, 4, 5], [15, 17, 2, 3],
              [(W + 1) / 2, (H + 1) / 2, W / 2, H / 2]],
    "expected": [[[1, 0, 0], [0, 1, 0]],
                 [[1 / 4, 0, 3 / 19], [0, 2 / 15, -11 / 29]],
                 [[3 / 20, 0, 13 / 19], [0, 1 / 15, -1 / 29]],
                 [[0.5, 0, 0], [0, 0.5, 0]]],
}

# ---- test/ApplyBoxTransform_test.lua:12-35 ----------------------------------------------
cases["apply_box_transform"] = {
    "src": "test/ApplyBoxTransform_test.lua:12-35", "tol": 1e-5,
    "boxes": [[10, 20, 30, 40], [-30, 50, 100, 200]],
    "trans": [[0.5, -0.2, 0, L(2)], [0, -0.5, L(0.5), L(1.1)]],
    "expected": [[25, 12, 30, 80], [-30, -50, 50, 220]],
}

# ---- test/MakeBoxes_test.lua:47-234 (tied to MakeAnchors+Reshape+ApplyBoxTransform by
#      test/MakeAnchors_test.lua:17-44).  expected[n][h][w][a] = (xc,yc,w,h), 0-based here.
log1_1, log2, log1_5, log0_5, log0_9 = 0.095310179804325, 0.69314718055995, 0.40546510810816, -0.69314718055995, -0.10536051565783


def mb(src, N, k, H, W, x0, y0, sx, sy, anchors, inputs, expected):
    # inputs: list of (n, a, h, w, [tx,ty,tw,th]) 0-based ; expected: list of (n, h, w, a, [xc,yc,w,h])
    return {"src": src, "N": N, "k": k, "H": H, "W": W, "x0": x0, "y0": y0, "sx": sx, "sy": sy,
            "anchors": anchors, "inputs": inputs, "expected": expected, "tol": 1e-4}


cases["make_boxes"] = [
    mb("test/MakeBoxes_test.lua:47-57", 1, 1, 1, 1, 1.5, 2.5, 1.0, 1.0, [[10], [20]], [],
       [(0, 0, 0, 0, [1.5, 2.5, 10, 20])]),
    mb("test/MakeBoxes_test.lua:60-75", 1, 3, 1, 1, 1.5, 2.5, 1.0, 1.0, [[10, 30, 100], [20, 40, 200]], [],
       [(0, 0, 0, 0, [1.5, 2.5, 10, 20]), (0, 0, 0, 1, [1.5, 2.5, 30, 40]), (0, 0, 0, 2, [1.5, 2.5, 100, 200])]),
    mb("test/MakeBoxes_test.lua:78-94", 1, 1, 2, 3, 1.5, 2.5, 1.0, 2.0, [[10], [20]], [],
       [(0, 0, 0, 0, [1.5, 2.5, 10, 20]), (0, 0, 1, 0, [2.5, 2.5, 10, 20]), (0, 0, 2, 0, [3.5, 2.5, 10, 20]),
        (0, 1, 0, 0, [1.5, 4.5, 10, 20]), (0, 1, 1, 0, [2.5, 4.5, 10, 20]), (0, 1, 2, 0, [3.5, 4.5, 10, 20])]),
    mb("test/MakeBoxes_test.lua:96-109", 2, 1, 1, 1, 2.0, 3.0, 10.0, 20.0, [[100], [200]], [],
       [(0, 0, 0, 0, [2.0, 3.0, 100, 200]), (1, 0, 0, 0, [2.0, 3.0, 100, 200])]),
    mb("test/MakeBoxes_test.lua:112-131", 2, 2, 1, 1, 2.0, 3.0, 1.0, 2.0, [[100, 10], [200, 20]],
       [(0, 0, 0, 0, [0.25, 0.1, 0, 0]), (0, 1, 0, 0, [0.1, 0.25, 0, 0]),
        (1, 0, 0, 0, [0, 0.05, 0, 0]), (1, 1, 0, 0, [0.05, 0, 0, 0])],
       [(0, 0, 0, 0, [27.0, 23.0, 100, 200]), (0, 0, 0, 1, [3.0, 8.0, 10, 20]),
        (1, 0, 0, 0, [2, 13, 100, 200]), (1, 0, 0, 1, [2.5, 3, 10, 20])]),
    mb("test/MakeBoxes_test.lua:134-157", 2, 2, 1, 1, 2.0, 3.0, 1.0, 2.0, [[100, 10], [200, 20]],
       [(0, 0, 0, 0, [0, 0, log1_1, 0]), (0, 1, 0, 0, [0, 0, 0, log2]),
        (1, 0, 0, 0, [0, 0, log1_5, log1_1]), (1, 1, 0, 0, [0, 0, log2, log1_5])],
       [(0, 0, 0, 0, [2, 3, 110, 200]), (0, 0, 0, 1, [2, 3, 10, 40]),
        (1, 0, 0, 0, [2, 3, 150, 220]), (1, 0, 0, 1, [2, 3, 20, 30])]),
    mb("test/MakeBoxes_test.lua:160-230", 2, 2, 2, 2, 2.0, 3.0, 1.0, 2.0, [[100, 10], [200, 20]],
       [(0, 0, 0, 0, [0, 0, 0, 0]), (0, 1, 0, 0, [0, 0, 0, log1_5]),
        (0, 0, 0, 1, [0, 0, log0_5, 0]), (0, 1, 0, 1, [0, 0, log2, log1_5]),
        (0, 0, 1, 0, [0, -0.02, 0, 0]), (0, 1, 1, 0, [0, 0.1, 0, log1_5]),
        (0, 0, 1, 1, [0, 0.1, log1_1, 0]), (0, 1, 1, 1, [0, 0.25, log1_5, log2]),
        (1, 0, 0, 0, [-0.05, 0, 0, 0]), (1, 1, 0, 0, [0.5, 0, 0, log0_5]),
        (1, 0, 0, 1, [0.1, 0, log1_1, 0]), (1, 1, 0, 1, [0.2, 0, log0_5, log1_5]),
        (1, 0, 1, 0, [-0.01, 0.1, 0, 0]), (1, 1, 1, 0, [1.1, 2.0, 0, log2]),
        (1, 0, 1, 1, [-1, 1, log0_9, 0]), (1, 1, 1, 1, [0.1, -0.2, log1_1, log0_9])],
       [(0, 0, 0, 0, [2, 3, 100, 200]), (0, 0, 0, 1, [2, 3, 10, 30]),
        (0, 0, 1, 0, [3, 3, 50, 200]), (0, 0, 1, 1, [3, 3, 20, 30]),
        (0, 1, 0, 0, [2, 1, 100, 200]), (0, 1, 0, 1, [2, 7, 10, 30]),
        (0, 1, 1, 0, [3, 25, 110, 200]), (0, 1, 1, 1, [3, 10, 15, 40]),
        (1, 0, 0, 0, [-3, 3, 100, 200]), (1, 0, 0, 1, [7, 3, 10, 10]),
        (1, 0, 1, 0, [13, 3, 110, 200]), (1, 0, 1, 1, [5, 3, 5, 30]),
        (1, 1, 0, 0, [1, 25, 100, 200]), (1, 1, 0, 1, [13, 45, 10, 40]),
        (1, 1, 1, 0, [-97, 205, 90, 200]), (1, 1, 1, 1, [4, 1, 11, 18])]),
]

# ---- test/ReshapeBoxFeatures_test.lua:33-58 : row order r = (a*H + y)*W + x ---------------
cases["reshape_consistency"] = {
    "src": "test/ReshapeBoxFeatures_test.lua:33-58", "N": 2, "k": 2, "D": 5, "H": 4, "W": 3,
    "anchors": [[10, 20], [20, 10]], "x0": 1, "y0": 1, "sx": 2, "sy": 2,
    # Lua 1-based: transforms[{2,{1,4},3,2}] = {10,10,0,0}; features[{2,{1,D},3,2}] = 100
    "n": 1, "a": 0, "y": 2, "x": 1, "trans": [10, 10, 0, 0], "feat_fill": 100,
    "expected_row_1based": 8, "expected_box": [103, 205, 10, 20],
}

# ---- test/LanguageModel_test.lua:135-160 (decodeSequence), :46-71 (getTarget) -------------
cases["decode_sequence"] = {
    "src": "test/LanguageModel_test.lua:135-160", "idx_to_token": ["a", "cat", "dog", "eating", "hungry"], "V": 5,
    "seq": [[1, 5, 2, 4, 1, 3, 6], [1, 3, 6, 0, 0, 0, 0], [2, 3, 1, 3, 2, 6, 0]],
    "expected": ["a hungry cat eating a dog", "a dog", "cat dog a dog cat"],
}

# ---- stale tests that do NOT agree with HEAD: documented divergence --
cases["documented_divergence"] = {
    "box_iou": {"src": "test/BoxIoU_test.lua:13-94 vs densecap/modules/BoxIoU.lua:40-73",
                "box1": [10, 10, 10, 10], "box2": [15, 15, 10, 10],
                "stale_test_expects": 25 / 175, "head_gives": 16 / 184},
}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")
with open(out, "w") as f:
    json.dump(cases, f, indent=1)
print("wrote", out)
