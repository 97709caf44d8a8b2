"""CPU-only checks of the drop-in boundary: the C-ABI library builds/loads, exports every symbol
that include/densecap_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from densecap_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "densecap_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dc_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header but not exported" % n
    # and the ctypes table binds exactly the header's functions
    assert sorted(_lib.SIGNATURES) == names


def test_config_default_matches_reference_defaults():
    lib = _lib.load()
    cfg = _lib.DcConfig()
    assert lib.dc_config_default(C.byref(cfg)) == 0
    assert list(cfg.widths) == [64, 128, 256, 512, 512]          # VGG-16
    assert cfg.rpn_filters == 256 and cfg.num_anchors == 12      # LocalizationLayer.lua:74,613-618
    assert cfg.fc_dim == 4096 and cfg.pool_hw == 7               # DenseCapModel.lua:66-67
    assert cfg.rnn_size == 512 and cfg.encoding_size == 512      # DenseCapModel.lua:42-43
    assert cfg.max_proposals == 1000 and cfg.seq_length == 15
    assert list(cfg.field_centers) == [8.5, 8.5, 16.0, 16.0]     # net_utils.lua:106-140
    k = cfg.num_anchors
    assert [cfg.anchors[i] for i in range(3)] == [45.0, 90.0, 64.0]
    assert [cfg.anchors[k + i] for i in range(3)] == [90.0, 45.0, 64.0]
    assert lib.dc_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    cfg = _lib.DcConfig()
    lib.dc_config_default(C.byref(cfg))
    h = C.c_void_p()
    rc = lib.dc_create(C.byref(cfg), 0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.dc_last_error(None)
    from densecap_b200.runtime import Context
    with pytest.raises(RuntimeError):
        Context()
    from densecap_b200 import box_utils
    with pytest.raises(TypeError):
        box_utils.xcycwh_to_x1y1x2y2(torch.zeros(3, 4))       # CPU tensor: refused, not emulated


def test_null_arguments_return_error_codes():
    lib = _lib.load()
    assert lib.dc_config_default(None) != 0
    assert lib.dc_set_image_size(None, 10, 10) != 0
    assert lib.dc_destroy(None) == 0


def test_synth_weights_shapes():
    from densecap_b200.synth import ModelDims, make_weights, feature_hw
    d = ModelDims(widths=(64, 64, 64, 64, 64), fc_dim=128, vocab_size=50, rnn_size=64, encoding_size=64, seq_length=4)
    w = make_weights(d, seed=0)
    assert w["conv1_1.weight"].shape == (64, 3, 3, 3)
    assert w["fc6.weight"].shape == (128, 64 * 49)
    assert w["lm.lstm.weight"].shape == (128, 256)
    assert w["lm.vocab.weight"].shape == (51, 64) and w["lm.lookup.weight"].shape == (52, 64)
    assert feature_hw(600, 720) == (38, 45) and feature_hw(600, 720, False) == (37, 45)
