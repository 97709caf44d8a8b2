"""Dense-captioning evaluator core: the reference's mAP computation without its METEOR bridge.
  iou_matrix / merge_boxes        -- densecap/box_utils.lua:563-612 (Lua numbers = float64 arithmetic on float32 box values)
  DenseCaptioningEvaluator        -- eval/eval_utils.lua:134-312: per image, ground-truth boxes overlapping >= 0.7 are merged
                                     (mean box, pooled reference captions), detections are matched greedily in decreasing
                                     confidence to the best-overlapping ground-truth box; evaluate() sweeps IoU thresholds
                                     {0.3..0.7} x language-score thresholds {-1, 0, 0.05..0.25} and reports max-interpolated AP
                                     at 101 recall points (mAP over the language thresholds, detection mAP for score > -1).
The language score (METEOR through an external Java process in the reference, eval_utils.score_captions) is a pluggable
callable `scorer(records) -> list of floats`; `exact_match_scorer` is the trivial stand-in used by the tests.
Host logic only (numpy); indices are 0-based here."""
from __future__ import annotations

import numpy as np


def xcycwh_to_x1y1x2y2(boxes):
    b = np.asarray(boxes, np.float32).reshape(-1, 4)
    hw = (b[:, 2] + np.float32(-1)) / np.float32(2)
    hh = (b[:, 3] + np.float32(-1)) / np.float32(2)
    return np.stack([hw * np.float32(-1) + b[:, 0], hh * np.float32(-1) + b[:, 1], hw + b[:, 0], hh + b[:, 1]], axis=1).astype(np.float32)


def _iou(bb, bb2):
    """Lua-number (double) arithmetic of box_utils.iou_matrix / the evaluator's matching loop; 0 when the boxes do not intersect."""
    iw = min(bb[2], bb2[2]) - max(bb[0], bb2[0]) + 1
    ih = min(bb[3], bb2[3]) - max(bb[1], bb2[1]) + 1
    if iw > 0 and ih > 0:
        ua = (bb[2] - bb[0] + 1) * (bb[3] - bb[1] + 1) + (bb2[2] - bb2[0] + 1) * (bb2[3] - bb2[1] + 1) - iw * ih
        return iw * ih / ua
    return 0.0


def iou_matrix(boxes_x1y1x2y2):
    b = np.asarray(boxes_x1y1x2y2, np.float32).astype(np.float64)
    n = b.shape[0]
    D = np.zeros((n, n))
    for i in range(n):
        D[i, i] = 1.0
        for j in range(i + 1, n):
            D[i, j] = D[j, i] = _iou(b[i], b[j])
    return D


def merge_boxes(boxes_x1y1x2y2, thr):
    """Groups of box indices (0-based, ascending inside a group), in the order the reference forms them: repeatedly take the box
    with the most neighbours at IoU >= thr (first such box on ties, torch.max), absorb it with all of them."""
    assert thr > 0
    D = iou_matrix(boxes_x1y1x2y2)
    groups = []
    while True:
        good = D >= thr
        good_sum = good.sum(axis=0)
        top = int(np.argmax(good_sum))
        if good_sum[top] == 0:
            break
        idx = np.nonzero(good[top])[0]
        groups.append(idx)
        D[idx, :] = 0
        D[:, idx] = 0
    return groups


def exact_match_scorer(records):
    """1.0 when the candidate equals one of its references, else 0.0 (a stand-in for METEOR with the same interface)."""
    return [1.0 if r["candidate"] in r["references"] else 0.0 for r in records]


class DenseCaptioningEvaluator:
    MIN_OVERLAPS = (0.3, 0.4, 0.5, 0.6, 0.7)
    MIN_SCORES = (-1, 0, 0.05, 0.1, 0.15, 0.2, 0.25)

    def __init__(self, scorer=exact_match_scorer, reference_quirks=True):
        """reference_quirks: a detection that overlaps NO ground-truth box leaves jmax = -1 in eval_utils.lua:169-205, and the
        following `used[jmax]` is a Torch7 tensor index of -1, i.e. the LAST element (TH wraps negative indices): the unmatched
        detection silently consumes the `used` flag of the image's last merged ground-truth box.  True reproduces that (parity with
        the reference's numbers); False gives the intended behaviour."""
        self.scorer = scorer
        self.reference_quirks = reference_quirks
        self.all_logprobs = []
        self.records = []
        self.n = 1
        self.npos = 0

    def addResult(self, logprobs, boxes, text, target_boxes, target_text):
        """logprobs (B,2) or (B,) (column 0 = positive class), boxes (B,4) and target_boxes (M,4) in xcycwh."""
        logprobs = np.asarray(logprobs, np.float64)
        lp = logprobs[:, 0] if logprobs.ndim == 2 else logprobs
        boxes = np.asarray(boxes, np.float32)
        target_boxes = np.asarray(target_boxes, np.float32)
        assert lp.shape[0] == boxes.shape[0] == len(text) and target_boxes.shape[0] == len(target_text) and boxes.ndim == 2
        b = xcycwh_to_x1y1x2y2(boxes).astype(np.float64)
        tb = xcycwh_to_x1y1x2y2(target_boxes)
        groups = merge_boxes(tb, 0.7)
        merged = np.stack([tb[g].astype(np.float64).mean(axis=0) for g in groups]) if groups else np.zeros((0, 4))
        merged_text = [[target_text[j] for j in g] for g in groups]
        order = np.argsort(-lp, kind="stable")
        used = np.zeros(len(groups), bool)
        for ii in order:
            ovmax, jmax = 0.0, -1
            for j in range(len(groups)):
                ov = _iou(b[ii], merged[j])
                if ov > ovmax:
                    ovmax, jmax = ov, j
            ok = 0
            flag = jmax if jmax >= 0 else (len(groups) - 1 if (self.reference_quirks and len(groups) > 0) else -1)
            if flag >= 0 and not used[flag]:
                used[flag] = True
                ok = 1
            self.records.append({"ok": ok, "ov": ovmax, "candidate": text[ii], "references": merged_text[jmax] if jmax >= 0 else [],
                                 "aligned": jmax >= 0, "imgid": self.n})
        self.n += 1
        self.npos += len(groups)
        self.all_logprobs.append(lp[order])

    def numAdded(self):
        return self.n - 1

    def evaluate(self):
        logprobs = np.concatenate(self.all_logprobs) if self.all_logprobs else np.zeros((0,))
        scores = list(self.scorer(self.records))
        ix = np.argsort(-logprobs, kind="stable")
        ap_results, det_results = {}, {}
        for min_overlap in self.MIN_OVERLAPS:
            for min_score in self.MIN_SCORES:
                n = ix.shape[0]
                tp = np.zeros(n); fp = np.zeros(n)
                for i, ii in enumerate(ix):
                    r = self.records[ii]
                    if r["aligned"] and r["ov"] >= min_overlap and r["ok"] == 1 and scores[ii] > min_score:
                        tp[i] = 1
                    else:
                        fp[i] = 1
                fp = np.cumsum(fp); tp = np.cumsum(tp)
                rec = tp / max(self.npos, 1)
                prec = tp / np.maximum(fp + tp, 1e-300)
                # `for t=0,1,0.01` (eval_utils.lua:288) ACCUMULATES the step in double precision: after 100 additions
                # t = 1.0000000000000007 > 1, so the reference averages over 100 thresholds (0 .. ~0.99), not 101, and the
                # thresholds are the accumulated values (e.g. 0.060000000000000005), which decide `rec >= t` on exact ties.
                ap, apn, t = 0.0, 0, 0.0
                while t <= 1.0:
                    m = rec >= t
                    ap += float(prec[m].max()) if m.any() else 0.0
                    apn += 1
                    t += 0.01
                ap /= apn
                if min_score == -1:
                    det_results["ov%s" % min_overlap] = ap
                else:
                    ap_results["ov%s_score%s" % (min_overlap, min_score)] = ap
        mean = lambda d: float(np.mean(list(d.values()))) if d else 0.0
        return {"map": mean(ap_results), "ap_breakdown": ap_results, "detmap": mean(det_results), "det_breakdown": det_results}
