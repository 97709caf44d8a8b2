/*
 * densecap_b200.h -- flat C ABI of libdensecap_b200.so
 *
 * B200-native (sm_100a) drop-in for the DenseCap test-time hot path of jcjohnson/densecap
 * (reference paths below are relative to that repository @ 7c32170):
 *
 *   VGG-16 conv trunk -> nn.LocalizationLayer (RPN heads, MakeAnchors / ReshapeBoxFeatures /
 *   ApplyBoxTransform, clip, objectness, NMS, BilinearRoiPooling) -> recognition head
 *   (fc6/fc7, objectness, box regression) -> nn.LanguageModel greedy LSTM decode -> final NMS.
 *
 * Conventions (identical to the reference at the boundary):
 *   - boxes are fp32 (xc, yc, w, h) in 1-based image pixel coordinates unless stated;
 *   - indices / tokens returned to the caller are 1-based int64 (Lua LongTensor);
 *   - all tensor pointers are DEVICE pointers unless the parameter name ends in _host;
 *   - every function returns 0 on success or a negative DC_E_* code, never throws/exits;
 *     dc_last_error() returns the message (the Lua/Python wrapper turns it into error()).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *     asynchronous with respect to the host unless stated.
 *
 * The header is plain C and can be pasted into a LuaJIT ffi.cdef.
 */
#ifndef DENSECAP_B200_H
#define DENSECAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DC_OK            0
#define DC_E_INVALID    -1   /* bad argument / shape                                   */
#define DC_E_PROTOCOL   -2   /* call-order violation (e.g. forward without setImageSize) */
#define DC_E_CUDA       -3   /* CUDA runtime / driver error (sticky per ctx)           */
#define DC_E_NOMEM      -4
#define DC_E_UNSUPPORTED -5
#define DC_E_NOTFOUND   -6   /* unknown weight name                                     */
#define DC_E_RANGE      -7   /* an activation left the fp16 range of the split hi+lo planes */

#define DC_MAX_ANCHORS 32

typedef struct dc_ctx dc_ctx;

/* Effective defaults = the reference's (DenseCapModel.lua:22-43,61-67;
 * LocalizationLayer.lua:64-88,613-618; net_utils.lua:106-140). */
typedef struct dc_config {
  int32_t widths[5];        /* VGG-16 stage widths, default 64,128,256,512,512              */
  int32_t rpn_filters;      /* default 256  (LocalizationLayer.lua:74)                       */
  int32_t num_anchors;      /* default 12   (LocalizationLayer.lua:613-618)                  */
  int32_t fc_dim;           /* default 4096 (DenseCapModel.lua:67)                           */
  int32_t pool_hw;          /* default 7    (DenseCapModel.lua:66)                           */
  int32_t rnn_size;         /* default 512  (DenseCapModel.lua:43)                           */
  int32_t encoding_size;    /* default 512  (DenseCapModel.lua:42)                           */
  int32_t vocab_size;       /* V; logits are V+1 wide, lookup has V+2 rows                   */
  int32_t seq_length;       /* default 15                                                    */
  int32_t pool_ceil;        /* 1 = loadcaffe ceil-mode max-pool (default)                    */
  int32_t max_batch;        /* images per dc_forward_test call (reference: 1)                */
  int32_t max_image_h, max_image_w;
  int32_t max_proposals;    /* capacity for num_proposals (default 1000)                     */
  int32_t reserved0;        /* must be 0 (was: engine selector; the library has ONE dense engine, tcgen05) */
  int32_t fuse_pool;        /* 1 = 2x2 ceil-mode max-pool fused into the conv epilogue        */
  float   anchors[2 * DC_MAX_ANCHORS]; /* (2,k) row-major: widths then heights               */
  float   field_centers[4]; /* x0, y0, sx, sy; default 8.5, 8.5, 16, 16                      */
} dc_config;

/* Result of one test-time forward (DenseCapModel:forward_test, DenseCapModel.lua:242-282,319-327).
 * All pointers are ctx-owned DEVICE buffers, valid until the next forward on this ctx
 * (same contract as the reference's reused self.output).  Per image b (0-based):
 *   rows [b*max_proposals, b*max_proposals + num_final[b]) are valid. */
typedef struct dc_result {
  int32_t batch;
  int32_t max_proposals;         /* row stride between images                                    */
  const int32_t* num_proposals;  /* (batch) K  : proposals kept by the RPN NMS                   */
  const int32_t* num_final;      /* (batch) K' : boxes kept by the final NMS                     */
  const float*   final_boxes;    /* (batch*P,4) xcycwh, NMS-filtered, score order   output[4]    */
  const float*   final_scores;   /* (batch*P)   objectness of those boxes           output[1]    */
  const int64_t* final_seq;      /* (batch*P,T) tokens in [1,V+1]                   output[5]    */
  const int64_t* final_idx;      /* (batch*P)   1-based row (among the K proposals) of each kept box */
  const float*   roi_boxes;      /* (batch*P,4) proposals, unfiltered               output[2]    */
  const float*   box_trans;      /* (batch*P,4) final box transforms, unfiltered    output[3]    */
  const float*   all_boxes;      /* (batch*P,4) final boxes before the final NMS                 */
  const float*   all_scores;     /* (batch*P)                                                    */
  const int64_t* all_seq;        /* (batch*P,T) tokens by proposal row; rows dropped by the final NMS are 0
                                    unless dc_set_decode_all(ctx, 1)                               */
  const float*   codes;          /* (batch*P,fc_dim) fc7 codes (recog_base output)               */
  const int64_t* proposal_idx;   /* (batch*P)   1-based row, among the clipped-valid RPN boxes, of each proposal
                                    (`idx` of LocalizationLayer.lua:322-338; dump_vars 'rpn_idx')    */
} dc_result;

/* ---- context ---------------------------------------------------------------------------- */
int dc_config_default(dc_config* cfg);
int dc_create(const dc_config* cfg, int device, dc_ctx** out);
int dc_destroy(dc_ctx* ctx);
const char* dc_last_error(const dc_ctx* ctx);      /* ctx may be NULL: last ctx-less error    */
int dc_version(void);
/* Debugging aid: synchronises the current device and pops the CUDA runtime's last error (0 = none; msg receives its text).  With
 * DENSECAP_B200_DEBUG=1 the Python binding calls it after every entry point and reports the first call that leaves an error behind. */
int dc_debug_peek_cuda_error(char* msg, int capacity);

/* Weights arrive in the reference's own layouts (host fp32) and are re-laid-out on the device:
 *   "convI_J.weight" (Cout,Cin,3,3) / ".bias";  "rpn.conv|box|score.weight/.bias";
 *   "fc6|fc7|obj|boxreg.weight" (out,in) / ".bias";  "lm.encoder.weight/.bias";
 *   "lm.lookup.weight" (V+2,E);  "lm.lstm.weight" ((E+H),4H) [i|f|o|g] / "lm.lstm.bias";
 *   "lm.vocab.weight" (V+1,H) / ".bias".
 * Replaces torch.load(checkpoint).model + model:convert() (run_model.lua:146-149). */
int dc_load_weights(dc_ctx* ctx, const char* name, const float* data_host,
                    const int64_t* shape, int ndim);

/* DenseCapModel:setTestArgs (DenseCapModel.lua:185-191) + LocalizationLayer:setTestArgs (:233-238) */
int dc_set_test_args(dc_ctx* ctx, float rpn_nms_thresh, float final_nms_thresh,
                     int num_proposals, int clip_boxes);
/* The reference decodes a caption for every proposal and then keeps only the rows selected by the final NMS
 * (DenseCapModel.lua:261-275).  By default this library runs the final NMS first and decodes only the surviving rows (rows are
 * independent, so output[5] is identical).  enable = 1 decodes every proposal, which fills dc_result.all_seq completely
 * (debugging / parity of the unfiltered LM output). */
int dc_set_decode_all(dc_ctx* ctx, int enable);
/* LanguageModel decode mode of the full forward: 0 = greedy LanguageModel:sample (default), n > 0 = LanguageModel:beamsearch with
 * beam_size n <= 32 (the reference dispatches on self.beam_size, LanguageModel.lua:129-132). */
int dc_set_beam_size(dc_ctx* ctx, int beam_size);
/* Range guard.  Dense activations travel as split-fp16 planes (hi + lo), exact only for |x| < 65504; every kernel that writes such
 * planes flags values outside that range (or NaN) in a sticky device word.  dc_check_range synchronises `stream`, returns DC_E_RANGE if
 * a forward since the last check overflowed (its results are invalid) and clears the flag.  The host-boundary calls
 * (dc_forward_test_host / _u8) check it themselves. */
int dc_check_range(dc_ctx* ctx, void* stream);
/* Accuracy / speed knob of the tensor-core kernels.  The tcgen05 accumulator adds every MMA result with truncation, which drifts by
 * ~0.5 ulp per accumulation; the kernels therefore close the TMEM accumulator every g k-blocks (64 K each) and add the partial sums in
 * fp32 registers with round-to-nearest.  g per tile width (64 / 128 / 256 output columns); 1 = closest to an fp32 sgemm, larger = fewer
 * accumulator sweeps.  The defaults are tuned on B200. */
int dc_set_accumulation_groups(dc_ctx* ctx, int kb_tile64, int kb_tile128, int kb_tile256);
/* LocalizationLayer:setImageSize (LocalizationLayer.lua:167-172): must precede each forward. */
int dc_set_image_size(dc_ctx* ctx, int image_h, int image_w);

/* DenseCapModel:forward (test mode).  img: (batch,3,H,W) fp32 NCHW, BGR, mean-subtracted
 * (run_model.lua:70-74).  dc_forward_test calls dc_set_image_size(H,W) itself exactly like
 * DenseCapModel:updateOutput (DenseCapModel.lua:245-246).
 * Concurrency: a context is used by one host thread at a time, and ONE forward runs per device at a time.  The caption decoder is a
 * persistent kernel whose CTAs (2 per 176 output columns, 144 for DenseCap's LM) wait for each other through device-side counters and
 * therefore need to be co-resident; forwards of two contexts interleaved on the same GPU could each hold part of the SMs.  The waits are
 * bounded (a violated assumption surfaces as DC_E_CUDA after ~2 s, never as a hung GPU).  Use one process / context per GPU and batch the
 * images (max_batch) instead -- which is also faster. */
int dc_forward_test(dc_ctx* ctx, const float* img, int batch, int image_h, int image_w,
                    dc_result* out, void* stream);

/* Same call with HOST buffers at the boundary (what a LuaJIT caller holding FloatTensors
 * would use): copies img_host H2D, runs, copies final results D2H into caller arrays
 * (each may be NULL), synchronises.  Capacity of out arrays: batch*max_proposals rows. */
int dc_forward_test_host(dc_ctx* ctx, const float* img_host, int batch, int image_h, int image_w,
                         int32_t* num_final_host, float* final_boxes_host, float* final_scores_host,
                         int64_t* final_seq_host);

/* SURVEY.md 8f rank 1 -- the step immediately before the path: the image preprocessing of run_model.lua:66-74 /
 * webcam/daemon.lua:66-78 without the resize (image.scale is third-party and stays with the caller): decoded uint8 RGB
 * HWC (batch,H,W,3) -> float/255 -> BGR -> *255 -> minus the VGG mean (103.939, 116.779, 123.68) -> fp32 NCHW.
 * dc_preprocess_u8: device pointers.  dc_forward_test_host_u8: dc_forward_test_host with the uint8 image at the
 * boundary (4x fewer host->device bytes; the preprocessing runs on the device). */
int dc_preprocess_u8(dc_ctx* ctx, const uint8_t* img_u8, int batch, int image_h, int image_w,
                     float* img_nchw, void* stream);
int dc_forward_test_host_u8(dc_ctx* ctx, const uint8_t* img_host_u8, int batch, int image_h, int image_w,
                            int32_t* num_final_host, float* final_boxes_host, float* final_scores_host,
                            int64_t* final_seq_host);
/* The same preprocessing WITH the resize of run_model.lua:69 / webcam/daemon.lua:70, image.scale(img, image_size): the larger side
 * becomes image_size (aspect kept, fractions truncated), bilinear mode = two separable fp32 passes (torch `image` rock, restated in
 * oracle.image_scale; third-party arithmetic, parity unpinned).  img_u8 (image_h, image_w, 3) DEVICE uint8 RGB; img_nchw must hold
 * 3 * image_size * image_size floats; the output size is returned through the HOST ints out_h / out_w. */
int dc_preprocess_u8_resize(dc_ctx* ctx, const uint8_t* img_u8, int image_h, int image_w, int image_size, float* img_nchw,
                            int* out_h, int* out_w, void* stream);
/* DenseCapModel:extractFeatures (DenseCapModel.lua:285-304): the network without the language model; final NMS of the final boxes
 * by objectness; boxes_xcycwh (K',4) and feats (K',fc_dim) = recog_base output of the kept rows in score order; num_out (1) int32.
 * Capacity max_proposals rows each; any output may be NULL. */
int dc_extract_features(dc_ctx* ctx, const float* img, int image_h, int image_w, float* boxes_xcycwh, float* feats,
                        int32_t* num_out, void* stream);

/* ---- stages of the path (each mirrors one reference entry point; used for per-stage parity) */

/* conv_net1+conv_net2 (DenseCapModel.lua:61-76).  feats_nchw out: (batch,C5,H',W') fp32. */
int dc_vgg_trunk_forward(dc_ctx* ctx, const float* img, int batch, int image_h, int image_w,
                         float* feats_nchw, void* stream);

/* nn.LocalizationLayer:_forward_test (LocalizationLayer.lua:250-363) on caller-provided
 * features (1,C,H',W') fp32 NCHW.  Requires dc_set_image_size first (DC_E_PROTOCOL otherwise,
 * message 'Must call setImageSize before each forward pass').
 * Outputs (any may be NULL): roi_features (K,C,HH,WW) fp32 BCHW; roi_boxes (K,4);
 * idx (K) 1-based int64 rows among the clipped-valid boxes; num_out (1) int32 = K. */
int dc_localization_forward_test(dc_ctx* ctx, const float* feats_nchw, int feat_h, int feat_w,
                                 float* roi_features, float* roi_boxes, int64_t* idx,
                                 int32_t* num_out, void* stream);

/* The second half of LocalizationLayer:_forward_test (LocalizationLayer.lua:262-338) on the RPN's raw outputs as the
 * reference's nn.Sequential RPN returns them: rpn_trans (1,4k,H',W') and rpn_scores (1,2k,H',W') fp32 NCHW (channel 4a+d /
 * 2a+d).  MakeAnchors + ReshapeBoxFeatures + ApplyBoxTransform + clip_boxes + objectness + box_utils.nms + the index gathers.
 * Every step is exact fp32 arithmetic in the reference's operation order, so idx is bit-exact for identical inputs.
 * Requires dc_set_image_size (consumed).  Outputs (any may be NULL): roi_boxes (K,4) xcycwh; idx (K) 1-based int64 rows
 * among the clipped-valid boxes, in pick order; num_out (1) int32 = K.  Capacity num_proposals rows. */
int dc_rpn_proposals(dc_ctx* ctx, const float* rpn_trans_nchw, const float* rpn_scores_nchw, int feat_h, int feat_w,
                     float* roi_boxes, int64_t* idx, int32_t* num_out, void* stream);

/* ---- training step of the LocalizationLayer (BASELINE configs[4]): forward + backward
 * Options of nn.LocalizationLayer used by _forward_train (LocalizationLayer.lua:79-87; train_opts.lua:18-42).
 * sampler_batch_size <= min(max_batch*max_proposals, 1024). */
int dc_set_train_args(dc_ctx* ctx, float mid_box_reg_weight, float mid_objectness_weight,
                      float box_reg_decay, float sampler_low_thresh, float sampler_high_thresh,
                      int sampler_batch_size, int train_remove_outbounds_boxes);
/* nn.BoxSampler minibatch selection (BoxSampler.lua:138-159).  mode 1 (default): uniform random like torch.multinomial -- positives
 * without replacement, negatives without replacement unless there are fewer than requested -- from a counter-based generator:
 * the same seed replays the same sequence of minibatches.  mode 0: the first P positives / first M negatives in index order, i.e.
 * the reference's debug_pos_sample_idx / debug_neg_sample_idx injection (what the parity tests feed the oracle). */
int dc_set_sampler(dc_ctx* ctx, int mode, uint64_t seed);
/* LocalizationLayer:setGroundTruth (LocalizationLayer.lua:181-186): gt_boxes (B1,4) xcycwh fp32, host or
 * device memory; must precede every training-time forward.  gt_labels stay with the caller: the layer
 * only indexes them with pos_target_idx (:455-457). */
int dc_set_ground_truth(dc_ctx* ctx, const float* gt_boxes, int num_gt, void* stream);
/* nn.LocalizationLayer:_forward_train (LocalizationLayer.lua:383-526): RPN on ALL k*H'*W' boxes,
 * nn.BoxSampler (modules/BoxSampler.lua:64-167) with the out-of-bounds exclusion, BoxSamplerHelper
 * gathers, ROI pooling of the S = sampler_batch_size sampled boxes (positives first), and the mid-level
 * losses (objectness cross-entropy, SmoothL1 on InvertBoxTransform targets with the |t|>10 mask hack,
 * box_reg_decay regulariser).  The minibatch is drawn as configured by dc_set_sampler (uniform random by default).
 * Errors: DC_E_PROTOCOL 'Must call setGroundTruth before training-time forward pass' /
 * 'Must call setImageSize before each forward pass' (same order as the reference's asserts).
 * Outputs (device, any may be NULL): roi_features (S,C,HH,WW); roi_boxes (S,4); pos_target_boxes (S,4;
 * first P rows valid); pos_target_idx / pos_input_idx / neg_input_idx 1-based int64 (S each; P / P / M
 * valid); num_pos_neg int32[4] = {total positives, total negatives, P, M}; losses fp32[5] =
 * {obj_loss_pos, obj_loss_neg, box_reg_loss, box_decay_loss, total}. */
int dc_localization_forward_train(dc_ctx* ctx, const float* feats_nchw, int feat_h, int feat_w,
                                  float* roi_features, float* roi_boxes, float* pos_target_boxes,
                                  int64_t* pos_target_idx, int64_t* pos_input_idx,
                                  int64_t* neg_input_idx, int32_t* num_pos_neg, float* losses,
                                  void* stream);

/* nn.LocalizationLayer:backward for the training path = updateGradInput (LocalizationLayer.lua:529-605: criterion
 * gradients scaled by the loss weights, BilinearRoiPooling backward [stnbhwd sampler / affine-grid backward +
 * BoxToAffine.lua:96-118], BoxSamplerHelper scatter [indexCopy, modules/BoxSamplerHelper.lua:142-182], ApplyBoxTransform
 * backward [:93-128], RegularizeLayer backward, ReshapeBoxFeatures backward, RPN convolutions backward) plus
 * accGradParameters of the RPN convolutions.  Must follow dc_localization_forward_train of the same input on this
 * context; dc_set_image_size / dc_set_ground_truth arm one forward AND one backward call each, like the reference's
 * _called_backward_size / _called_backward_gt flags (DC_E_PROTOCOL with the reference's assert messages otherwise).
 * grad_roi_features (S,C,HH,WW), grad_roi_boxes (S,4) = gradOutput[1], gradOutput[2];
 * grad_feats_nchw (1,C,H',W') = gradInput.  All device pointers. */
int dc_localization_backward_train(dc_ctx* ctx, const float* feats_nchw, int feat_h, int feat_w,
                                   const float* grad_roi_features, const float* grad_roi_boxes,
                                   float* grad_feats_nchw, void* stream);
/* nn.Module:zeroGradParameters / read-out of gradWeight, gradBias by the reference's parameter names
 * ("rpn.conv.weight" (F,C,3,3), "rpn.conv.bias", "rpn.box.weight" (4k,F), "rpn.box.bias", "rpn.score.weight"
 * (2k,F), "rpn.score.bias").  Gradients accumulate over backward calls until zeroed.  out: host or device. */
int dc_zero_grad_parameters(dc_ctx* ctx, void* stream);
int dc_get_grad_parameters(dc_ctx* ctx, const char* name, float* out, int64_t numel, void* stream);

/* recog_base + objectness + box_reg + ApplyBoxTransform (DenseCapModel.lua:87-100,127-162).
 * roi_features (K,C,HH,WW) fp32 BCHW, roi_boxes (K,4).  Outputs (K,fc_dim),(K),(K,4),(K,4). */
int dc_recog_forward(dc_ctx* ctx, const float* roi_features, const float* roi_boxes, int K,
                     float* codes, float* obj_scores, float* box_trans, float* final_boxes,
                     void* stream);

/* LanguageModel:sample (LanguageModel.lua:293-348), greedy argmax.  codes (K,fc_dim).
 * seq (K,T) int64 tokens in [1,V+1].  logits (optional, may be NULL): (K,T,V+1) fp32. */
int dc_lm_sample(dc_ctx* ctx, const float* codes, int K, int64_t* seq, float* logits, void* stream);
/* LanguageModel:beamsearch(image_vectors, beam_size) (LanguageModel.lua:170-290): per image vector the beam_size best partial
 * captions are extended token by token (log-softmax scores, END-token masking by multiplication, state re-indexing), the best beam
 * is returned.  beam_size <= 32 (<= 0: the reference default 20).  seq (K,T) int64. */
int dc_lm_beamsearch(dc_ctx* ctx, const float* codes, int K, int beam_size, int64_t* seq, void* stream);

/* ---- leaf operators (stateless; one per reference module) ------------------------------- */

/* nn.MakeAnchors:updateOutput (modules/MakeAnchors.lua:40-67): out (N,4k,H,W). anchors_host (2,k). */
int dc_make_anchors(float* out, int N, int H, int W, int k, float x0, float y0, float sx, float sy,
                    const float* anchors_host, void* stream);
/* nn.ReshapeBoxFeatures:updateOutput (modules/ReshapeBoxFeatures.lua:24-33): (N,kD,H,W)->(N,kHW,D) */
int dc_reshape_box_features(const float* in, float* out, int N, int k, int D, int H, int W, void* stream);
/* nn.ApplyBoxTransform:updateOutput (modules/ApplyBoxTransform.lua:63-90): n boxes */
int dc_apply_box_transform(const float* boxes, const float* trans, float* out, int n, void* stream);
/* nn.InvertBoxTransform:updateOutput (modules/InvertBoxTransform.lua:36-60) */
int dc_invert_box_transform(const float* anchors, const float* targets, float* out, int n, void* stream);
/* box_utils.xcycwh_to_x1y1x2y2 (box_utils.lua:270-298) / x1y1x2y2_to_xcycwh (:382-410) */
int dc_xcycwh_to_x1y1x2y2(const float* in, float* out, int n, void* stream);
int dc_x1y1x2y2_to_xcycwh(const float* in, float* out, int n, void* stream);
/* box_utils.clip_boxes(boxes, bounds, 'xcycwh') (box_utils.lua:486-523): valid (n) uint8 */
int dc_clip_boxes(const float* boxes, float* out, uint8_t* valid, int n,
                  float x_min, float y_min, float x_max, float y_max, void* stream);
/* objectness probability (LocalizationLayer.lua:305-308): scores (n,2) -> prob (n) */
int dc_objectness_prob(const float* scores, float* prob, int n, void* stream);
/* box_utils.nms (box_utils.lua:154-256).  boxes_scores (n,5) = x1,y1,x2,y2,score.
 * max_boxes <= 0 means no cap.  idx_out (>= min(n,max_boxes)) 1-based int64 in
 * decreasing-score order; count_out (1) int32.  Uses the ctx workspace. */
int dc_nms(dc_ctx* ctx, const float* boxes_scores, int n, float thresh, int max_boxes,
           int64_t* idx_out, int32_t* count_out, void* stream);
/* nn.BoxToAffine:updateOutput (modules/BoxToAffine.lua:69-93): (B,4) -> (B,2,3) */
int dc_box_to_affine(const float* boxes, float* theta, int B, int image_h, int image_w, void* stream);
/* nn.BilinearRoiPooling:updateOutput (modules/BilinearRoiPooling.lua:35-67,84-91) =
 * BoxToAffine + stnbhwd AffineGridGeneratorBHWD + BatchBilinearSamplerBHWD
 * (modules/BatchBilinearSamplerBHWD.lua:104-122).  feats (C,H',W') fp32 CHW, boxes (B,4).
 * out_layout 0: (B,C,HH,WW) as the reference returns; 1: (B,HH,WW,C) (native BHWD). */
int dc_bilinear_roi_pool(dc_ctx* ctx, const float* feats_chw, int C, int feat_h, int feat_w,
                         const float* boxes, int B, int image_h, int image_w, int HH, int WW,
                         float* out, int out_layout, void* stream);
/* stnbhwd nn.AffineGridGeneratorBHWD(HH,WW):updateOutput (call site modules/BilinearRoiPooling.lua:52): theta (B,2,3) ->
 * grids (B,HH,WW,2), last dimension (y, x) in [-1,1] normalised coordinates. */
int dc_affine_grid_generator(const float* theta, int B, int HH, int WW, float* grids, void* stream);
/* nn.BatchBilinearSamplerBHWD:updateOutput (modules/BatchBilinearSamplerBHWD.lua:104-122) = stnbhwd BilinearSamplerBHWD_updateOutput
 * on ONE image shared by all grids (the reference expands it with a 0-stride batch dimension, :114): feats_hwc (H',W',C) fp32,
 * grids (B,HH,WW,2) caller-supplied -> out (B,HH,WW,C).  HH*WW <= 64. */
int dc_batch_bilinear_sampler(dc_ctx* ctx, const float* feats_hwc, int feat_h, int feat_w, int C, const float* grids, int B,
                              int HH, int WW, float* out, void* stream);
/* nn.BoxIoU:updateOutput (modules/BoxIoU.lua:40-73): (B1,4),(B2,4) -> (B1,B2) */
int dc_box_iou(const float* box1, const float* box2, float* out, int B1, int B2, void* stream);

/* ---- dense kernels exposed for parity tests / micro-benchmarks --------------------------
 * y = act(x W^T + b), x (M,K) fp32, W (N,K) fp32 (nn.Linear layout), y (M,N) fp32.
 * relu: 0/1.  K must be a multiple of 64. */
int dc_linear(dc_ctx* ctx, const float* x, const float* w, const float* b, float* y,
              int M, int N, int K, int relu, void* stream);
/* y = relu(conv3x3(x) + b) [+ optional 2x2 ceil max-pool], x (N,Cin,H,W) fp32 NCHW,
 * w (Cout,Cin,3,3), y (N,Cout,H,W) or pooled (pool 1: stand-alone pooling kernel, 2: fused epilogue).
 * Cin must be a multiple of 64. */
int dc_conv3x3_relu(dc_ctx* ctx, const float* x, const float* w, const float* b, float* y,
                    int N, int Cin, int Cout, int H, int W, int pool, void* stream);

/* ---- instrumentation -------------------------------------------------------------------- */
/* Per-phase device times of the last dc_forward_test (milliseconds, cudaEvent based), in the
 * spirit of LocalizationLayer stats.times (LocalizationLayer.lua:219-230).  Enabling timing
 * inserts events only (no host sync until dc_get_times).  names_out: static strings.
 * enable = 2 additionally marks every convolution layer of the trunk ("conv_trunk" is then conv1_1 alone, followed by "conv1_2" ...). */
int dc_enable_timing(dc_ctx* ctx, int enable);
int dc_get_times(dc_ctx* ctx, const char** names_out, float* ms_out, int capacity, int* count_out);
/* Number of kernels launched by this library on ctx since creation (for bench accounting). */
int64_t dc_kernel_launches(const dc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DENSECAP_B200_H */
