"""Deterministic synthetic weights / inputs.

There are no checkpoints or datasets in this environment, so the bench and the parity
tests run random-init weights of the reference architecture, generated on the host from
one seeded numpy generator and handed byte-for-byte to both the CUDA path and the oracle.
Layouts are the reference's own (Torch7): conv (Cout,Cin,kH,kW), Linear (out,in),
torch-rnn LSTM ((D+H),4H) with gate blocks [i|f|o|g].
"""
from __future__ import annotations

import numpy as np

VGG16_CONVS = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3",
               "conv4_1", "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3"]
VGG16_STAGE = {"conv1": 0, "conv2": 1, "conv3": 2, "conv4": 3, "conv5": 4}
VGG_MEAN_BGR = np.array([103.939, 116.779, 123.68], np.float32)   # run_model.lua:72


class ModelDims:
    """Architecture dimensions; defaults are the reference's VGG-16 DenseCap model
    (densecap/DenseCapModel.lua:61-67,42-43; LocalizationLayer.lua:74,613-618)."""

    def __init__(self, widths=(64, 128, 256, 512, 512), rpn_filters=256, num_anchors=12,
                 fc_dim=4096, pool_hw=7, rnn_size=512, encoding_size=512, vocab_size=10497,
                 seq_length=15):
        self.widths = tuple(widths)
        self.rpn_filters = rpn_filters
        self.num_anchors = num_anchors
        self.fc_dim = fc_dim
        self.pool_hw = pool_hw
        self.rnn_size = rnn_size
        self.encoding_size = encoding_size
        self.vocab_size = vocab_size
        self.seq_length = seq_length

    @property
    def feat_dim(self):
        return self.widths[4]

    def conv_shapes(self):
        cin = 3
        out = []
        for name in VGG16_CONVS:
            cout = self.widths[VGG16_STAGE[name[:5]]]
            out.append((name, cout, cin))
            cin = cout
        return out


def feature_hw(h, w, pool_ceil=True):
    """Spatial size after the 4 2x2/2 max-pools of VGG-16 conv trunk (pool5 excluded)."""
    for _ in range(4):
        if pool_ceil:
            h, w = (h + 1) // 2, (w + 1) // 2
        else:
            h, w = h // 2, w // 2
    return h, w


def make_weights(dims: ModelDims, seed=0, box_variant="A"):
    """box_variant 'A': reference default init (RPN box conv and final box_reg zero ->
    proposals are clipped anchors; LocalizationLayer.lua:645-646, DenseCapModel.lua:98-99).
    'B': small random box heads so that exp()/transform paths are exercised.
    'C': 'B' with conv1_1 scaled by 1/64, i.e. the input normalised to unit scale.  He-init VGG preserves the variance of its
         +-128 input, so with 'A'/'B' the conv5_3 features are O(100) where a trained VGG's are O(1): at 720x600 the RPN score
         differences then have std ~20 and ~20 % of the 20,520 objectness probabilities round to exactly 1.0f (mass ties whose
         membership flips on last-bit noise), and 'B' box deltas reach exp(40).  'C' keeps every stage in the numeric range of a
         trained model (scores spread over (0,1), |tw| < 1): the well-conditioned full-size parity workload."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    w = {}

    def normal(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    def uniform(shape, a):
        return rng.uniform(-a, a, size=shape).astype(np.float32)

    calibrated = (box_variant == "C")
    if calibrated:
        box_variant = "B"
    for name, cout, cin in dims.conv_shapes():
        w[name + ".weight"] = normal((cout, cin, 3, 3), np.sqrt(2.0 / (9 * cin)))
        w[name + ".bias"] = np.zeros((cout,), np.float32)
    if calibrated:
        w["conv1_1.weight"] = (w["conv1_1.weight"] * np.float32(1.0 / 64.0)).astype(np.float32)
    C = dims.feat_dim
    k = dims.num_anchors
    F = dims.rpn_filters
    w["rpn.conv.weight"] = normal((F, C, 3, 3), 0.01)
    w["rpn.conv.bias"] = np.zeros((F,), np.float32)
    w["rpn.box.weight"] = np.zeros((4 * k, F, 1, 1), np.float32) if box_variant == "A" else normal((4 * k, F, 1, 1), 0.01)
    w["rpn.box.bias"] = np.zeros((4 * k,), np.float32)
    w["rpn.score.weight"] = normal((2 * k, F, 1, 1), 0.01)
    w["rpn.score.bias"] = np.zeros((2 * k,), np.float32)
    fan6 = C * dims.pool_hw * dims.pool_hw
    w["fc6.weight"] = normal((dims.fc_dim, fan6), np.sqrt(2.0 / fan6))
    w["fc6.bias"] = np.zeros((dims.fc_dim,), np.float32)
    w["fc7.weight"] = normal((dims.fc_dim, dims.fc_dim), np.sqrt(2.0 / dims.fc_dim))
    w["fc7.bias"] = np.zeros((dims.fc_dim,), np.float32)
    w["obj.weight"] = normal((1, dims.fc_dim), 0.01)
    w["obj.bias"] = np.zeros((1,), np.float32)
    w["boxreg.weight"] = np.zeros((4, dims.fc_dim), np.float32) if box_variant == "A" else normal((4, dims.fc_dim), 0.001)
    w["boxreg.bias"] = np.zeros((4,), np.float32)
    E, H, V = dims.encoding_size, dims.rnn_size, dims.vocab_size
    w["lm.encoder.weight"] = uniform((E, dims.fc_dim), 1.0 / np.sqrt(dims.fc_dim))
    w["lm.encoder.bias"] = uniform((E,), 1.0 / np.sqrt(dims.fc_dim))
    w["lm.lookup.weight"] = normal((V + 2, E), 1.0)
    w["lm.lstm.weight"] = normal((E + H, 4 * H), 1.0 / np.sqrt(E + H))
    b = np.zeros((4 * H,), np.float32)
    b[H:2 * H] = 1.0                                    # torch-rnn forget-gate bias init
    w["lm.lstm.bias"] = b
    w["lm.vocab.weight"] = uniform((V + 1, H), 1.0 / np.sqrt(H))
    w["lm.vocab.bias"] = uniform((V + 1,), 1.0 / np.sqrt(H))
    return w


def make_image(h, w, seed=0):
    """fp32 (1,3,H,W) BGR, pixel ~ U[0,255) minus the VGG mean (mirrors run_model.lua:70-74)."""
    rng = np.random.default_rng(np.random.PCG64(1000 + seed))
    img = rng.uniform(0.0, 255.0, size=(1, 3, h, w)).astype(np.float32)
    img -= VGG_MEAN_BGR.reshape(1, 3, 1, 1)
    return img
