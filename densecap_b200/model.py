"""Host-side mirrors of the reference's composite modules for the test-time path:
nn.LocalizationLayer (densecap/LocalizationLayer.lua), nn.LanguageModel (densecap/LanguageModel.lua)
and DenseCapModel (densecap/DenseCapModel.lua).  Same method names, argument meaning and protocol
errors; all compute happens in libdensecap_b200.so on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .runtime import Context, ptr, stream_ptr, _check_cuda_f32
from .synth import ModelDims, feature_hw


class LocalizationLayer:
    """nn.LocalizationLayer: test-time forward (LocalizationLayer.lua:250-363) and training-time forward (:383-526).

    forward(cnn_features (1,C,H',W')) -> [roi_features (K,C,HH,WW), roi_boxes (K,4), gt_boxes_sample, gt_labels_sample];
    the last two are empty at test time.  After a training forward self.stats['losses'] holds the mid-level losses."""

    def __init__(self, ctx: Context, opt=None):
        self.ctx = ctx
        self.train = False
        self.image_height = None
        self.image_width = None
        self.gt_labels = None
        self.stats = {"losses": {}, "vars": {}}
        opt = dict(opt or {})
        # defaults of LocalizationLayer.lua:79-87 (mid_* weights are mandatory there; train_opts.lua:32-35 values here)
        self.opt = {"sampler_batch_size": min(opt.get("sampler_batch_size", 256), ctx.max_proposals, 1024),
                    "sampler_high_thresh": opt.get("sampler_high_thresh", 0.7), "sampler_low_thresh": opt.get("sampler_low_thresh", 0.5),
                    "train_remove_outbounds_boxes": opt.get("train_remove_outbounds_boxes", 1),
                    "mid_box_reg_weight": opt.get("mid_box_reg_weight", 0.05), "mid_objectness_weight": opt.get("mid_objectness_weight", 0.1),
                    "box_reg_decay": opt.get("box_reg_decay", 0.0)}
        # torch.multinomial minibatch draws by default (BoxSampler.lua:138-152); {"sampler_mode": "deterministic"} selects the first
        # P / first M candidates in index order (the reference's debug_*_sample_idx hook)
        self.ctx.set_sampler(opt.get("sampler_mode", "random"), opt.get("sampler_seed", 0))
        self._push_train_args()
        self.setTestArgs()

    def _push_train_args(self):
        o = self.opt
        self.ctx.check(self.ctx._lib.dc_set_train_args(self.ctx.handle, o["mid_box_reg_weight"], o["mid_objectness_weight"], o["box_reg_decay"],
                                                      o["sampler_low_thresh"], o["sampler_high_thresh"], int(o["sampler_batch_size"]),
                                                      int(o["train_remove_outbounds_boxes"])))

    def setGroundTruth(self, gt_boxes, gt_labels):                      # LocalizationLayer.lua:181-186
        b = _check_cuda_f32(gt_boxes)
        assert b.dim() == 3 and b.shape[0] == 1 and b.shape[2] == 4, "gt_boxes must have shape (N, B1, 4)"
        assert gt_labels.dim() == 3 and gt_labels.shape[0] == 1 and gt_labels.shape[1] == b.shape[1], "gt_labels must have shape (N, B1, L)"
        self.gt_labels = gt_labels
        self.ctx.check(self.ctx._lib.dc_set_ground_truth(self.ctx.handle, ptr(b), int(b.shape[1]), stream_ptr()))

    def training(self):
        self.train = True

    def zeroGradParameters(self):
        self.ctx.check(self.ctx._lib.dc_zero_grad_parameters(self.ctx.handle, stream_ptr()))

    def backward(self, input, gradOutput):
        """nn.Module:backward at training time (LocalizationLayer.lua:529-605 + accGradParameters of the RPN):
        gradOutput = [grad_roi_features (S,C,HH,WW), grad_roi_boxes (S,4), ...]; returns gradInput (1,C,H',W')."""
        assert self.train, "can only call updateGradInput in training mode"
        f = _check_cuda_f32(input)
        g_feat = _check_cuda_f32(gradOutput[0])
        g_box = _check_cuda_f32(gradOutput[1])
        _, C_, fh, fw = f.shape
        S = int(self.opt["sampler_batch_size"])
        assert tuple(g_feat.shape) == (S, C_, self.ctx.dims.pool_hw, self.ctx.dims.pool_hw) and tuple(g_box.shape) == (S, 4)
        self.gradInput = torch.empty_like(f)
        self.ctx.check(self.ctx._lib.dc_localization_backward_train(self.ctx.handle, ptr(f), fh, fw, ptr(g_feat), ptr(g_box), ptr(self.gradInput),
                                                                   stream_ptr()))
        return self.gradInput

    def gradParameters(self):
        """dict name -> gradWeight / gradBias tensor (reference shapes) accumulated since zeroGradParameters."""
        d = self.ctx.dims
        k, F, C5 = d.num_anchors, d.rpn_filters, d.widths[4]
        shapes = {"rpn.conv.weight": (F, C5, 3, 3), "rpn.conv.bias": (F,), "rpn.box.weight": (4 * k, F, 1, 1), "rpn.box.bias": (4 * k,),
                  "rpn.score.weight": (2 * k, F, 1, 1), "rpn.score.bias": (2 * k,)}
        out = {}
        for name, shp in shapes.items():
            t = torch.empty(shp, dtype=torch.float32, device="cuda")
            self.ctx.check(self.ctx._lib.dc_get_grad_parameters(self.ctx.handle, name.encode(), ptr(t), t.numel(), stream_ptr()))
            out[name] = t
        return out

    def _forward_train(self, f):
        dims = self.ctx.dims
        _, C_, fh, fw = f.shape
        S = int(self.opt["sampler_batch_size"])
        dev = f.device
        roi = torch.empty((S, C_, dims.pool_hw, dims.pool_hw), dtype=torch.float32, device=dev)
        boxes = torch.empty((S, 4), dtype=torch.float32, device=dev)
        tgt_boxes = torch.empty((S, 4), dtype=torch.float32, device=dev)
        pt = torch.zeros((S,), dtype=torch.int64, device=dev)
        pi = torch.zeros((S,), dtype=torch.int64, device=dev)
        ni = torch.zeros((S,), dtype=torch.int64, device=dev)
        counts = torch.zeros((4,), dtype=torch.int32, device=dev)
        losses = torch.zeros((5,), dtype=torch.float32, device=dev)
        # raises DenseCapError(DC_E_PROTOCOL, 'Must call setGroundTruth ...' / 'Must call setImageSize ...') like the asserts at :385-393
        self.ctx.check(self.ctx._lib.dc_localization_forward_train(self.ctx.handle, ptr(f), fh, fw, ptr(roi), ptr(boxes), ptr(tgt_boxes), ptr(pt),
                                                                  ptr(pi), ptr(ni), ptr(counts), ptr(losses), stream_ptr()))
        tot_pos, tot_neg, P, M = [int(v) for v in counts.tolist()]
        lv = losses.tolist()
        self.stats = {"losses": {"obj_loss_pos": lv[0], "obj_loss_neg": lv[1], "box_reg_loss": lv[2], "box_decay_loss": lv[3],
                                 "total_loss": lv[4]},
                      "vars": {"num_pos": P, "num_neg": M, "total_pos": tot_pos, "total_neg": tot_neg}}
        self.pos_input_idx, self.pos_target_idx, self.neg_input_idx = pi[:P], pt[:P], ni[:M]
        labels = self.gt_labels[0].to(dev)[self.pos_target_idx - 1] if self.gt_labels is not None else torch.empty((0,), device=dev)
        self.output = [roi, boxes, tgt_boxes[:P], labels]
        return self.output

    def setImageSize(self, image_height, image_width):                 # LocalizationLayer.lua:167-172
        self.image_height, self.image_width = int(image_height), int(image_width)
        self.ctx.check(self.ctx._lib.dc_set_image_size(self.ctx.handle, self.image_height, self.image_width))

    def setTestArgs(self, args=None):                                   # LocalizationLayer.lua:233-238
        args = args or {}
        self.test_clip_boxes = args.get("clip_boxes", True)
        self.test_nms_thresh = args.get("nms_thresh", 0.7)
        self.test_max_proposals = min(args.get("max_proposals", 300), self.ctx.max_proposals)   # reference default 300
        # final_nms_thresh belongs to DenseCapModel (DenseCapModel.lua:185-191); the layer's setTestArgs touches only nms_thresh,
        # max_proposals and clip_boxes (LocalizationLayer.lua:233-238), so the context's current final threshold is passed through
        self.ctx.set_test_args(self.test_nms_thresh, None, self.test_max_proposals, self.test_clip_boxes)

    def evaluate(self):
        self.train = False

    def forward(self, input):
        f = _check_cuda_f32(input)
        assert f.dim() == 4 and f.shape[0] == 1, "must have 1 image per batch"
        if self.train:
            return self._forward_train(f)
        dims = self.ctx.dims
        _, C_, fh, fw = f.shape
        K = self.ctx.max_proposals if self.test_max_proposals == -1 else self.test_max_proposals
        roi = torch.empty((K, C_, dims.pool_hw, dims.pool_hw), dtype=torch.float32, device=f.device)
        boxes = torch.empty((K, 4), dtype=torch.float32, device=f.device)
        idx = torch.empty((K,), dtype=torch.int64, device=f.device)
        n = torch.zeros((1,), dtype=torch.int32, device=f.device)
        # raises DenseCapError(DC_E_PROTOCOL, 'Must call setImageSize before each forward pass') like the assert at :259
        self.ctx.check(self.ctx._lib.dc_localization_forward_test(self.ctx.handle, ptr(f), fh, fw, ptr(roi), ptr(boxes), ptr(idx), ptr(n),
                                                                 stream_ptr()))
        k = int(n.item())
        self.last_idx = idx[:k]
        empty = torch.empty((0,), dtype=torch.float32, device=f.device)
        self.output = [roi[:k], boxes[:k], empty, empty]
        return self.output


class LanguageModel:
    """nn.LanguageModel, greedy sampling path (LanguageModel.lua:293-348) + decodeSequence (:86-103)."""

    def __init__(self, ctx: Context, idx_to_token=None):
        self.ctx = ctx
        d = ctx.dims
        self.vocab_size = d.vocab_size
        self.seq_length = d.seq_length
        self.START_TOKEN = d.vocab_size + 1
        self.END_TOKEN = d.vocab_size + 1
        self.NULL_TOKEN = d.vocab_size + 2
        self.idx_to_token = idx_to_token or {}
        self.beam_size = None                   # like the reference: set it to run beam search in forward()

    def beamsearch(self, image_vectors, beam_size=20):
        """LanguageModel:beamsearch (LanguageModel.lua:170-290)."""
        v = _check_cuda_f32(image_vectors)
        N = v.shape[0]
        seq = torch.zeros((N, self.seq_length), dtype=torch.int64, device=v.device)
        self.ctx.check(self.ctx._lib.dc_lm_beamsearch(self.ctx.handle, ptr(v), N, int(beam_size or 20), ptr(seq), stream_ptr()))
        self.output = seq
        return seq

    def sample(self, image_vectors, return_logits=False):
        v = _check_cuda_f32(image_vectors)
        N = v.shape[0]
        T = self.seq_length
        seq = torch.zeros((N, T), dtype=torch.int64, device=v.device)
        logits = torch.empty((N, T, self.vocab_size + 1), dtype=torch.float32, device=v.device) if return_logits else None
        self.ctx.check(self.ctx._lib.dc_lm_sample(self.ctx.handle, ptr(v), N, ptr(seq), ptr(logits), stream_ptr()))
        self.output = seq
        return (seq, logits) if return_logits else seq

    def forward(self, input):
        image_vectors, gt_sequence = input
        if gt_sequence is not None and gt_sequence.numel() > 0:
            raise NotImplementedError("training-time LanguageModel forward is outside the test-time hot path")
        if self.beam_size is not None:          # LanguageModel.lua:129-132
            return self.beamsearch(image_vectors, self.beam_size)
        return self.sample(image_vectors)

    def decodeSequence(self, seq):
        captions = []
        for row in seq.tolist():
            words = []
            for idx in row:
                if idx == self.END_TOKEN or idx == 0:
                    break
                words.append(self.idx_to_token.get(idx, str(idx)) if isinstance(self.idx_to_token, dict) else self.idx_to_token[idx - 1])
            captions.append(" ".join(words))
        return captions


class DenseCapModel:
    """DenseCapModel test-time API (DenseCapModel.lua:185-191, 242-282, 319-327)."""

    def __init__(self, opt=None, weights=None, device=0, max_batch=1, max_image_hw=(600, 720), fuse_pool=True):
        opt = dict(opt or {})
        self.opt = opt
        dims = opt.get("dims") or ModelDims(vocab_size=opt.get("vocab_size", 10497), seq_length=opt.get("seq_length", 15),
                                            rnn_size=opt.get("rnn_size", 512), encoding_size=opt.get("rnn_encoding_size", 512))
        self.dims = dims
        self.opt.setdefault("final_nms_thresh", 0.3)
        self.ctx = Context(dims, device=device, max_batch=max_batch, max_image_hw=max_image_hw,
                           max_proposals=opt.get("max_proposals", 1000), fuse_pool=fuse_pool,
                           pool_ceil=opt.get("pool_ceil", True), anchors=opt.get("anchors"), field_centers=opt.get("field_centers"))
        self.train = False
        self.nets = {"localization_layer": LocalizationLayer(self.ctx),
                     "language_model": LanguageModel(self.ctx, opt.get("idx_to_token"))}
        self.setTestArgs({})
        if weights is not None:
            self.ctx.load_weights(weights)

    @classmethod
    def from_checkpoint(cls, path, device=0, max_batch=1, max_image_hw=(600, 720), max_proposals=1000):
        """`torch.load(opt.checkpoint).model` (run_model.lua:146-149) for a Torch7 .t7 checkpoint written by train.lua:184:
        parses the file (densecap_b200/t7.py), maps the module tree onto dc_load_weights names and sizes the context from
        the tensor shapes.  Returns a ready DenseCapModel with its idx_to_token vocabulary."""
        from . import t7
        weights, info = t7.extract_densecap_weights(t7.load(path))
        widths = (weights["conv1_2.weight"].shape[0], weights["conv2_2.weight"].shape[0], weights["conv3_3.weight"].shape[0],
                  weights["conv4_3.weight"].shape[0], weights["conv5_3.weight"].shape[0])
        E = weights["lm.encoder.weight"].shape[0]
        H = weights["lm.lstm.weight"].shape[1] // 4
        dims = ModelDims(widths=widths, rpn_filters=weights["rpn.conv.weight"].shape[0], num_anchors=weights["rpn.score.weight"].shape[0] // 2,
                         fc_dim=weights["fc7.weight"].shape[0], rnn_size=H, encoding_size=E, vocab_size=weights["lm.vocab.weight"].shape[0] - 1,
                         seq_length=int(info.get("seq_length") or 15))
        tok = info.get("idx_to_token")
        if isinstance(tok, dict):
            tok = {int(k): v for k, v in tok.items()}
        opt = {"dims": dims, "max_proposals": max_proposals, "idx_to_token": tok}
        # model state stored in the checkpoint overrides the library defaults: MakeAnchors' anchor table and field centres, pooling mode
        for key in ("anchors", "field_centers", "pool_ceil"):
            if key in info:
                opt[key] = info[key]
        return cls(opt, weights=weights, device=device, max_batch=max_batch, max_image_hw=max_image_hw)

    def evaluate(self):
        self.train = False

    def setTestArgs(self, kwargs):
        self._rpn_nms = kwargs.get("rpn_nms_thresh", 0.7)
        self.opt["final_nms_thresh"] = kwargs.get("final_nms_thresh", 0.3)
        self._num_proposals = min(kwargs.get("num_proposals", 1000), self.ctx.max_proposals)   # reference default 1000
        self.ctx.set_test_args(self._rpn_nms, self.opt["final_nms_thresh"], self._num_proposals, True)

    def forward_device(self, input):
        """Runs the whole test-time forward; returns a dict of torch CUDA tensors that alias ctx-owned
        buffers (valid until the next forward -- same contract as the reference's self.output)."""
        x = _check_cuda_f32(input)
        assert x.dim() == 4 and x.shape[1] == 3, "input must be (N,3,H,W)"
        N, _, H, W = x.shape
        res = _lib.DcResult()
        self.ctx.check(self.ctx._lib.dc_forward_test(self.ctx.handle, ptr(x), N, H, W, C.byref(res), stream_ptr()))
        return _wrap_result(res, self.dims, x.device)

    def forward(self, input):
        """DenseCapModel:updateOutput for one image (1,3,H,W): list of the 7 outputs
        (DenseCapModel.lua:223-241) with [1],[4],[5] NMS-filtered (:272-274)."""
        assert input.dim() == 4 and input.shape[0] == 1 and input.shape[1] == 3
        r = self.forward_device(input)
        K = int(r["num_proposals"][0].item())
        Kf = int(r["num_final"][0].item())
        empty = torch.empty((0,), device=input.device)
        self.output = [r["final_scores"][:Kf].reshape(Kf, 1), r["roi_boxes"][:K], r["box_trans"][:K], r["final_boxes"][:Kf],
                       r["final_seq"][:Kf], empty, empty]
        self._last = r
        return self.output

    def extractFeatures(self, input):
        """DenseCapModel:extractFeatures (DenseCapModel.lua:285-304): boxes (K',4) xcycwh and recog_base features (K',fc_dim) of the
        boxes kept by the final NMS, in score order; the language model does not run."""
        x = _check_cuda_f32(input)
        assert x.dim() == 4 and x.shape[0] == 1 and x.shape[1] == 3
        P = self.ctx.max_proposals
        boxes = torch.empty((P, 4), dtype=torch.float32, device=x.device)
        feats = torch.empty((P, self.dims.fc_dim), dtype=torch.float32, device=x.device)
        n = torch.zeros((1,), dtype=torch.int32, device=x.device)
        self.ctx.check(self.ctx._lib.dc_extract_features(self.ctx.handle, ptr(x), x.shape[2], x.shape[3], ptr(boxes), ptr(feats), ptr(n),
                                                         stream_ptr()))
        k = int(n.item())
        return boxes[:k], feats[:k]

    def forward_test(self, input):
        """DenseCapModel:forward_test (DenseCapModel.lua:319-327)."""
        self.evaluate()
        out = self.forward(input)
        captions = self.nets["language_model"].decodeSequence(out[4])
        return out[3], out[0], captions


def preprocess_image(ctx, img_u8, image_size=None):
    """run_model.lua:66-74 on the device: uint8 RGB (N,H,W,3) or (H,W,3) CUDA tensor -> (N,3,H,W) fp32 BGR, mean-subtracted -- what
    DenseCapModel:forward_test expects.  image_size: image.scale(img, image_size) first (one image; larger side -> image_size)."""
    if not isinstance(img_u8, torch.Tensor) or not img_u8.is_cuda or img_u8.dtype != torch.uint8:
        raise TypeError("preprocess_image takes a CUDA uint8 tensor")
    x = img_u8.contiguous()
    if image_size is not None:
        if x.dim() == 4:
            assert x.shape[0] == 1, "the resizing path takes one image"
            x = x[0]
        H, W, C_ = x.shape
        assert C_ == 3
        out = torch.empty((3 * int(image_size) * int(image_size),), dtype=torch.float32, device=x.device)
        oh, ow = C.c_int(0), C.c_int(0)
        ctx.check(ctx._lib.dc_preprocess_u8_resize(ctx.handle, ptr(x), H, W, int(image_size), ptr(out), C.byref(oh), C.byref(ow), stream_ptr()))
        return out[:3 * oh.value * ow.value].view(1, 3, oh.value, ow.value)
    if x.dim() == 3:
        x = x[None]
    N, H, W, C_ = x.shape
    assert C_ == 3
    out = torch.empty((N, 3, H, W), dtype=torch.float32, device=x.device)
    ctx.check(ctx._lib.dc_preprocess_u8(ctx.handle, ptr(x), N, H, W, ptr(out), stream_ptr()))
    return out


def _dev_tensor(addr, shape, dtype, device):
    """Zero-copy torch view of a ctx-owned device buffer through the CUDA array interface."""
    n = int(np.prod(shape))
    if n == 0 or not addr:
        return torch.empty(shape, dtype=dtype, device=device)
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8"}[dtype]

    class _Holder:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(addr), False), "version": 2}

    return torch.as_tensor(_Holder(), device=device)


def _wrap_result(res, dims, device):
    B, P, T = res.batch, res.max_proposals, dims.seq_length
    rows = B * P
    return {
        "num_proposals": _dev_tensor(res.num_proposals, (B,), torch.int32, device),
        "num_final": _dev_tensor(res.num_final, (B,), torch.int32, device),
        "final_boxes": _dev_tensor(res.final_boxes, (rows, 4), torch.float32, device),
        "final_scores": _dev_tensor(res.final_scores, (rows,), torch.float32, device),
        "final_seq": _dev_tensor(res.final_seq, (rows, T), torch.int64, device),
        "final_idx": _dev_tensor(res.final_idx, (rows,), torch.int64, device),
        "roi_boxes": _dev_tensor(res.roi_boxes, (rows, 4), torch.float32, device),
        "box_trans": _dev_tensor(res.box_trans, (rows, 4), torch.float32, device),
        "all_boxes": _dev_tensor(res.all_boxes, (rows, 4), torch.float32, device),
        "all_scores": _dev_tensor(res.all_scores, (rows,), torch.float32, device),
        "all_seq": _dev_tensor(res.all_seq, (rows, T), torch.int64, device),
        "codes": _dev_tensor(res.codes, (rows, dims.fc_dim), torch.float32, device),
        "proposal_idx": _dev_tensor(res.proposal_idx, (rows,), torch.int64, device),
        "max_proposals": P,
    }
