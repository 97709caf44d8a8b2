"""ctypes binding of libdensecap_b200.so -- the same symbols a LuaJIT ffi.cdef would bind
(include/densecap_b200.h).  Loading fails loudly when the library is missing; there is no
Python/CPU fallback for any operator."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

DC_MAX_ANCHORS = 32

c_f = C.c_float
c_i = C.c_int
c_vp = C.c_void_p
c_i64 = C.c_int64
c_i32 = C.c_int32


class DcConfig(C.Structure):
    _fields_ = [("widths", c_i32 * 5), ("rpn_filters", c_i32), ("num_anchors", c_i32), ("fc_dim", c_i32),
                ("pool_hw", c_i32), ("rnn_size", c_i32), ("encoding_size", c_i32), ("vocab_size", c_i32),
                ("seq_length", c_i32), ("pool_ceil", c_i32), ("max_batch", c_i32), ("max_image_h", c_i32),
                ("max_image_w", c_i32), ("max_proposals", c_i32), ("reserved0", c_i32), ("fuse_pool", c_i32),
                ("anchors", c_f * (2 * DC_MAX_ANCHORS)), ("field_centers", c_f * 4)]


class DcResult(C.Structure):
    _fields_ = [("batch", c_i32), ("max_proposals", c_i32), ("num_proposals", c_vp), ("num_final", c_vp),
                ("final_boxes", c_vp), ("final_scores", c_vp), ("final_seq", c_vp), ("final_idx", c_vp),
                ("roi_boxes", c_vp), ("box_trans", c_vp), ("all_boxes", c_vp), ("all_scores", c_vp),
                ("all_seq", c_vp), ("codes", c_vp), ("proposal_idx", c_vp)]


# name -> (restype, argtypes): exactly the declarations of include/densecap_b200.h
SIGNATURES = {
    "dc_config_default": (c_i, [C.POINTER(DcConfig)]),
    "dc_create": (c_i, [C.POINTER(DcConfig), c_i, C.POINTER(c_vp)]),
    "dc_destroy": (c_i, [c_vp]),
    "dc_last_error": (C.c_char_p, [c_vp]),
    "dc_version": (c_i, []),
    "dc_load_weights": (c_i, [c_vp, C.c_char_p, c_vp, C.POINTER(c_i64), c_i]),
    "dc_set_test_args": (c_i, [c_vp, c_f, c_f, c_i, c_i]),
    "dc_set_decode_all": (c_i, [c_vp, c_i]),
    "dc_set_beam_size": (c_i, [c_vp, c_i]),
    "dc_check_range": (c_i, [c_vp, c_vp]),
    "dc_set_accumulation_groups": (c_i, [c_vp, c_i, c_i, c_i]),
    "dc_debug_peek_cuda_error": (c_i, [C.c_char_p, c_i]),
    "dc_set_sampler": (c_i, [c_vp, c_i, C.c_uint64]),
    "dc_set_image_size": (c_i, [c_vp, c_i, c_i]),
    "dc_forward_test": (c_i, [c_vp, c_vp, c_i, c_i, c_i, C.POINTER(DcResult), c_vp]),
    "dc_forward_test_host": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_vp]),
    "dc_preprocess_u8": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp, c_vp]),
    "dc_forward_test_host_u8": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_vp]),
    "dc_preprocess_u8_resize": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp, C.POINTER(c_i), C.POINTER(c_i), c_vp]),
    "dc_extract_features": (c_i, [c_vp, c_vp, c_i, c_i, c_vp, c_vp, c_vp, c_vp]),
    "dc_vgg_trunk_forward": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp, c_vp]),
    "dc_localization_forward_test": (c_i, [c_vp, c_vp, c_i, c_i, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dc_rpn_proposals": (c_i, [c_vp, c_vp, c_vp, c_i, c_i, c_vp, c_vp, c_vp, c_vp]),
    "dc_set_train_args": (c_i, [c_vp, c_f, c_f, c_f, c_f, c_f, c_i, c_i]),
    "dc_set_ground_truth": (c_i, [c_vp, c_vp, c_i, c_vp]),
    "dc_localization_forward_train": (c_i, [c_vp, c_vp, c_i, c_i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dc_localization_backward_train": (c_i, [c_vp, c_vp, c_i, c_i, c_vp, c_vp, c_vp, c_vp]),
    "dc_zero_grad_parameters": (c_i, [c_vp, c_vp]),
    "dc_get_grad_parameters": (c_i, [c_vp, C.c_char_p, c_vp, C.c_int64, c_vp]),
    "dc_recog_forward": (c_i, [c_vp, c_vp, c_vp, c_i, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dc_lm_sample": (c_i, [c_vp, c_vp, c_i, c_vp, c_vp, c_vp]),
    "dc_lm_beamsearch": (c_i, [c_vp, c_vp, c_i, c_i, c_vp, c_vp]),
    "dc_affine_grid_generator": (c_i, [c_vp, c_i, c_i, c_i, c_vp, c_vp]),
    "dc_batch_bilinear_sampler": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp, c_i, c_i, c_i, c_vp, c_vp]),
    "dc_make_anchors": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_vp, c_vp]),
    "dc_reshape_box_features": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "dc_apply_box_transform": (c_i, [c_vp, c_vp, c_vp, c_i, c_vp]),
    "dc_invert_box_transform": (c_i, [c_vp, c_vp, c_vp, c_i, c_vp]),
    "dc_xcycwh_to_x1y1x2y2": (c_i, [c_vp, c_vp, c_i, c_vp]),
    "dc_x1y1x2y2_to_xcycwh": (c_i, [c_vp, c_vp, c_i, c_vp]),
    "dc_clip_boxes": (c_i, [c_vp, c_vp, c_vp, c_i, c_f, c_f, c_f, c_f, c_vp]),
    "dc_objectness_prob": (c_i, [c_vp, c_vp, c_i, c_vp]),
    "dc_nms": (c_i, [c_vp, c_vp, c_i, c_f, c_i, c_vp, c_vp, c_vp]),
    "dc_box_to_affine": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp]),
    "dc_bilinear_roi_pool": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp, c_i, c_i, c_i, c_i, c_i, c_vp, c_i, c_vp]),
    "dc_box_iou": (c_i, [c_vp, c_vp, c_vp, c_i, c_i, c_vp]),
    "dc_linear": (c_i, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_i, c_vp]),
    "dc_conv3x3_relu": (c_i, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "dc_enable_timing": (c_i, [c_vp, c_i]),
    "dc_get_times": (c_i, [c_vp, C.POINTER(C.c_char_p), C.POINTER(c_f), c_i, C.POINTER(c_i)]),
    "dc_kernel_launches": (c_i64, [c_vp]),
}


class DenseCapError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("densecap_b200 error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


def lib_path():
    # DENSECAP_B200_LIB: development hook to load a variant build of the same library (see build.build_library(defines=...))
    return os.environ.get("DENSECAP_B200_LIB") or _build.LIB_PATH


def load(build_if_missing=True):
    """Returns the ctypes handle of libdensecap_b200.so (built on demand when nvcc is present)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if build_if_missing and path == _build.LIB_PATH and _build.is_stale() and _build._nvcc() is not None:
        _build.build_library()
    if not os.path.exists(path):
        raise ImportError("libdensecap_b200.so not found at %s and could not be built; "
                          "the CUDA extension is required (no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("DENSECAP_B200_DEBUG"):
        lib = _DebugLib(lib)
    _LIB = lib
    return lib


class _DebugLib:
    """DENSECAP_B200_DEBUG=1: after every entry point, synchronise and report a CUDA error the call left unreported."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name in ("dc_debug_peek_cuda_error", "dc_last_error"):
            return fn
        peek = self._lib.dc_debug_peek_cuda_error

        def wrapped(*a):
            rc = fn(*a)
            buf = C.create_string_buffer(256)
            if peek(buf, 256) != 0:
                import sys
                print("DENSECAP_B200_DEBUG: %s returned %r but left CUDA error: %s" % (name, rc, buf.value.decode()), file=sys.stderr, flush=True)
            return rc
        return wrapped


def check(rc, ctx=None):
    if rc != 0:
        msg = load().dc_last_error(ctx)
        raise DenseCapError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")
