"""Torch7 serialisation reader/writer and the DenseCap checkpoint -> dc_load_weights mapping.
No Torch7 is available here, so the reader is checked (a) against a hand-assembled byte stream written out from the format
description of torch7's File.lua, and (b) by round trips through the writer on a checkpoint with the reference's module tree."""
import struct

import numpy as np
import pytest

from densecap_b200 import t7
from densecap_b200.synth import ModelDims, make_weights


def i32(v):
    return struct.pack("<i", v)


def i64(v):
    return struct.pack("<q", v)


def s(x):
    return i32(len(x)) + x.encode()


def test_reader_on_hand_assembled_stream():
    # { ["a"] = 2.5, [1] = true, ["t"] = FloatTensor(2,3) transposed view of a 3x2 storage, ["t2"] = same tensor again (shared) }
    storage = np.arange(6, dtype=np.float32)
    tensor = (i32(4) + i32(2) + s("V 1") + s("torch.FloatTensor") + i32(2) + i64(2) + i64(3) + i64(1) + i64(2) + i64(1)
              + i32(4) + i32(3) + s("V 1") + s("torch.FloatStorage") + i64(6) + storage.tobytes())
    blob = (i32(3) + i32(1) + i32(4)
            + i32(2) + s("a") + i32(1) + struct.pack("<d", 2.5)
            + i32(1) + struct.pack("<d", 1.0) + i32(5) + i32(1)
            + i32(2) + s("t") + tensor
            + i32(2) + s("t2") + i32(4) + i32(2))
    out = t7.loads(blob)
    assert out["a"] == 2.5 and out[1] is True
    assert out["t"].dtype == np.float32 and out["t"].tolist() == [[0, 2, 4], [1, 3, 5]]      # sizes (2,3), strides (1,2)
    assert out["t2"] is out["t"]                                                              # reference sharing by index


def test_round_trip_values_and_sharing():
    shared = {"x": 1, "y": "z"}
    obj = {"list": [1, 2.5, "three", False, None], "shared_a": shared, "shared_b": shared,
           "mod": t7.T7Object("nn.Linear", {"weight": np.random.RandomState(0).randn(3, 4).astype(np.float32), "bias": np.zeros(3, np.float32)}),
           "long": np.arange(5, dtype=np.int64), "dbl": np.linspace(0, 1, 4)}
    back = t7.loads(t7.dumps(obj))
    assert back["list"] == {1: 1, 2: 2.5, 3: "three", 4: False, 5: None}
    assert back["shared_a"] is back["shared_b"] and back["shared_a"] == {"x": 1, "y": "z"}
    assert back["mod"].classname == "nn.Linear" and np.array_equal(back["mod"]["weight"], obj["mod"]["weight"])
    assert back["long"].dtype == np.int64 and back["long"].tolist() == [0, 1, 2, 3, 4]
    assert back["dbl"].dtype == np.float64 and np.array_equal(back["dbl"], obj["dbl"])


fake_checkpoint = t7.densecap_checkpoint_tree


def test_densecap_checkpoint_round_trip(tmp_path):
    dims = ModelDims(widths=(8, 16, 16, 16, 16), rpn_filters=8, fc_dim=32, rnn_size=16, encoding_size=16, vocab_size=20, seq_length=4)
    w = make_weights(dims, seed=1, box_variant="B")
    if w["obj.weight"].ndim == 1:
        w = dict(w); w["obj.weight"] = w["obj.weight"].reshape(1, -1)
    path = tmp_path / "ckpt.t7"
    t7.save(fake_checkpoint(dims, w), str(path))
    got, info = t7.extract_densecap_weights(t7.load(str(path)))
    assert info["vocab_size"] == dims.vocab_size and info["idx_to_token"][3] == "w2"
    assert set(got) == set(w), sorted(set(got) ^ set(w))
    for k in w:
        assert got[k].dtype == np.float32
        assert np.array_equal(got[k].reshape(-1), np.asarray(w[k], np.float32).reshape(-1)), k
    assert got["conv1_1.weight"].shape == (8, 3, 3, 3) and got["rpn.box.weight"].shape == (4 * dims.num_anchors, 8, 1, 1)


def test_unsupported_and_truncated_streams():
    with pytest.raises(ValueError):
        t7.loads(i32(6) + i32(1))              # a serialised Lua function
    with pytest.raises(Exception):
        t7.loads(i32(2) + i32(10) + b"abc")    # truncated string
    with pytest.raises(ValueError):
        t7.loads(i32(42))


def test_results_json_matches_run_model_format(tmp_path):
    """run_model.lua:79,89-95,182-188: boxes leave as (x, y, w, h) = xcycwh_to_xywh, one entry per detection, wrapped with opt."""
    import json
    from densecap_b200 import results as R
    from oracle import densecap_oracle as O
    boxes = np.array([[10.5, 20.0, 8.0, 6.0], [300.0, 200.0, 101.0, 51.0]], np.float32)
    r = R.result_to_json(boxes, np.array([[0.5], [-1.25]], np.float32), ["a cat", "the sky"], img_name="elephant.jpg")
    assert np.array_equal(np.asarray(r["boxes"], np.float32), O.xcycwh_to_xywh(boxes))
    assert r["boxes"][0] == [7.0, 17.5, 8.0, 6.0] and r["scores"] == [0.5, -1.25] and r["captions"] == ["a cat", "the sky"]
    doc = R.write_results([r], {"image_size": 720, "num_proposals": 1000}, str(tmp_path / "results.json"))
    back = json.load(open(tmp_path / "results.json"))
    assert back == doc and back["results"][0]["img_name"] == "elephant.jpg" and set(back) == {"results", "opt"}
    with pytest.raises(ValueError):
        R.result_to_json(boxes, [0.5], ["x", "y"])


def test_checkpoint_model_state_is_extracted():
    """MakeAnchors' anchor table / field centres and the pooling mode stored in the module tree are reported (they were
    silently replaced by library defaults)."""
    from densecap_b200.synth import ModelDims, make_weights
    dims = ModelDims(widths=(8, 8, 8, 8, 8), rpn_filters=8, fc_dim=16, rnn_size=8, encoding_size=8, vocab_size=10, seq_length=3)
    w = make_weights(dims, seed=1)
    tree = t7.densecap_checkpoint_tree(dims, w)
    _, info = t7.extract_densecap_weights(tree)
    assert info["pool_ceil"] is True and "anchors" not in info          # the synthetic tree carries no MakeAnchors state
    loc = tree["model"].fields["nets"]["localization_layer"]
    ma = [m for m in t7._walk_all(loc, []) if m.classname == "nn.MakeAnchors"][0]
    anchors = np.arange(2 * dims.num_anchors, dtype=np.float32).reshape(2, -1) + 10
    ma.fields.update({"anchors": anchors, "x0": 7.5, "y0": 7.5, "sx": 16.0, "sy": 16.0})
    for m in t7._walk_all(tree["model"].fields["nets"]["conv_net1"], []):
        if m.classname == "nn.SpatialMaxPooling":
            m.fields["ceil_mode"] = False
    for m in t7._walk_all(tree["model"].fields["nets"]["conv_net2"], []):
        if m.classname == "nn.SpatialMaxPooling":
            m.fields["ceil_mode"] = False
    blob_path = None
    import io, tempfile, os
    with tempfile.TemporaryDirectory() as d:
        blob_path = os.path.join(d, "m.t7")
        t7.save(tree, blob_path)
        _, info2 = t7.extract_densecap_weights(t7.load(blob_path))
    assert np.array_equal(info2["anchors"], anchors)
    assert info2["field_centers"] == (7.5, 7.5, 16.0, 16.0)
    assert info2["pool_ceil"] is False
