"""The step after the path (serialisation part): the `results.json` document of run_model.lua.
  result_to_json   -- run_model.lua:89-95  {boxes (x, y, w, h), scores, captions} of one image
  write_results    -- run_model.lua:171-188 {"results": [... + img_name], "opt": {...}} written like utils.write_json
Boxes arrive as the model returns them, (xc, yc, w, h) in 1-based pixel coordinates, and are converted with
box_utils.xcycwh_to_xywh (box_utils.lua:311-380 composition) exactly as run_model.lua:79 does.  Host logic only."""
from __future__ import annotations

import json

import numpy as np


def xcycwh_to_xywh(boxes):
    """box_utils.xcycwh_to_xywh = x1y1x2y2_to_xywh(xcycwh_to_x1y1x2y2(boxes)) in the reference's fp32 operation order."""
    b = np.asarray(boxes, np.float32).reshape(-1, 4)
    half_w = (b[:, 2] + np.float32(-1)) / np.float32(2)
    half_h = (b[:, 3] + np.float32(-1)) / np.float32(2)
    x1 = half_w * np.float32(-1) + b[:, 0]
    x2 = half_w + b[:, 0]
    y1 = half_h * np.float32(-1) + b[:, 1]
    y2 = half_h + b[:, 1]
    return np.stack([x1, y1, (x2 - x1) + np.float32(1), (y2 - y1) + np.float32(1)], axis=1).astype(np.float32)


def result_to_json(boxes_xcycwh, scores, captions, img_name=None):
    boxes = xcycwh_to_xywh(np.asarray(boxes_xcycwh, np.float32))
    out = {"boxes": [[float(v) for v in row] for row in boxes],
           "scores": [float(v) for v in np.asarray(scores, np.float32).reshape(-1)],
           "captions": list(captions)}
    if len(out["scores"]) != len(out["boxes"]) or len(out["captions"]) != len(out["boxes"]):
        raise ValueError("result_to_json: boxes, scores and captions must have the same length")
    if img_name is not None:
        out["img_name"] = img_name
    return out


def write_results(results, opt, path):
    """results: list of result_to_json(...) dicts (each with img_name).  Writes {"results": ..., "opt": ...}."""
    doc = {"results": list(results), "opt": dict(opt or {})}
    with open(path, "w") as f:
        json.dump(doc, f)
    return doc
