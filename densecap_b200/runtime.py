"""Host-side plumbing: the dc_ctx handle, weight upload, torch tensor <-> raw pointer helpers.
PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .synth import ModelDims

_DEFAULT_CTX = {}


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_cuda_f32(t):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("densecap_b200 operators take CUDA tensors (there is no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Context:
    """Owns one dc_ctx (weights + workspace arena) on one GPU.  One Context per GPU; single-threaded."""

    def __init__(self, dims: ModelDims = None, device=0, max_batch=1, max_image_hw=(600, 720), max_proposals=1000,
                 fuse_pool=True, pool_ceil=True, anchors=None, field_centers=None):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("densecap_b200 needs a CUDA device (B200 / sm_100a); no CPU fallback exists")
        dims = dims or ModelDims()
        self.dims = dims
        self.device = int(device)
        cfg = _lib.DcConfig()
        _lib.check(lib.dc_config_default(C.byref(cfg)))
        for i, w in enumerate(dims.widths):
            cfg.widths[i] = w
        cfg.rpn_filters = dims.rpn_filters
        cfg.num_anchors = dims.num_anchors
        cfg.fc_dim = dims.fc_dim
        cfg.pool_hw = dims.pool_hw
        cfg.rnn_size = dims.rnn_size
        cfg.encoding_size = dims.encoding_size
        cfg.vocab_size = dims.vocab_size
        cfg.seq_length = dims.seq_length
        cfg.pool_ceil = 1 if pool_ceil else 0
        cfg.max_batch = max_batch
        cfg.max_image_h, cfg.max_image_w = max_image_hw
        cfg.max_proposals = max_proposals
        cfg.fuse_pool = 1 if fuse_pool else 0
        if anchors is not None:
            a = np.asarray(anchors, np.float32)
            assert a.shape == (2, dims.num_anchors)
            for i, v in enumerate(a.reshape(-1)):
                cfg.anchors[i] = float(v)
        if field_centers is not None:
            for i, v in enumerate(field_centers):
                cfg.field_centers[i] = float(v)
        self.cfg = cfg
        self.max_proposals = max_proposals
        self.max_batch = max_batch
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.init()
            _lib.check(lib.dc_create(C.byref(cfg), self.device, C.byref(h)))
        self.handle = h
        self._lib = lib

    def close(self):
        if getattr(self, "handle", None):
            self._lib.dc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        _lib.check(rc, self.handle)

    def load_weights(self, weights: dict):
        """weights: name -> numpy fp32 array in the reference layout (see include/densecap_b200.h)."""
        for name, arr in weights.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            self.check(self._lib.dc_load_weights(self.handle, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim))

    def set_test_args(self, rpn_nms_thresh=0.7, final_nms_thresh=0.3, num_proposals=1000, clip_boxes=True):
        """final_nms_thresh=None keeps the value of the last call (the LocalizationLayer does not own that option)."""
        if final_nms_thresh is None:
            final_nms_thresh = getattr(self, "_final_nms_thresh", 0.3)
        self._final_nms_thresh = float(final_nms_thresh)
        self.check(self._lib.dc_set_test_args(self.handle, float(rpn_nms_thresh), float(final_nms_thresh), int(num_proposals),
                                              1 if clip_boxes else 0))

    def set_decode_all(self, on=True):
        """Decode a caption for every proposal (fills `all_seq`) instead of only for the rows kept by the final NMS."""
        self.check(self._lib.dc_set_decode_all(self.handle, 1 if on else 0))

    def set_beam_size(self, beam_size):
        """0 = greedy LanguageModel:sample, n > 0 = LanguageModel:beamsearch(beam_size n) in the full forward."""
        self.check(self._lib.dc_set_beam_size(self.handle, int(beam_size or 0)))

    def set_sampler(self, mode="random", seed=0):
        """nn.BoxSampler minibatch draw: 'random' (torch.multinomial-like, seeded) or 'deterministic' (first P / first M)."""
        self.check(self._lib.dc_set_sampler(self.handle, 1 if mode == "random" else 0, int(seed)))

    def enable_timing(self, on=True):
        self.check(self._lib.dc_enable_timing(self.handle, 1 if on else 0))

    def get_times(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = C.c_int(0)
        self.check(self._lib.dc_get_times(self.handle, names, ms, 64, C.byref(n)))
        return [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    def kernel_launches(self):
        return int(self._lib.dc_kernel_launches(self.handle))


def default_context(device=0, min_boxes=0):
    """A small lazily created context for the stateless-looking operators that need workspace (nms, roi pool)."""
    key = int(device)
    ctx = _DEFAULT_CTX.get(key)
    if ctx is None or ctx._cap_boxes < min_boxes:
        if ctx is not None:
            ctx.close()
        k = 12
        cells = max(2048, (min_boxes + k - 1) // k)
        side = int(np.ceil(np.sqrt(cells)))
        dims = ModelDims(widths=(64, 64, 64, 64, 512), fc_dim=64, vocab_size=64, rnn_size=64, encoding_size=64)
        ctx = Context(dims, device=device, max_image_hw=(side * 16, side * 16), max_proposals=64)
        ctx._cap_boxes = k * side * side
        _DEFAULT_CTX[key] = ctx
    return ctx
