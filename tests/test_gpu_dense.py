"""GPU parity of the dense tcgen05 split-fp16 kernels
against an fp64 torch reference of the same op: y = act(x W^T + b) and 3x3 conv + bias + ReLU
(+ 2x2 ceil max-pool).  Tolerance: 1e-4 relative to the output scale (BASELINE north_star);
the split-fp16 scheme is expected to land around 1e-6."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from densecap_b200.runtime import Context
    from densecap_b200.synth import ModelDims
    d = ModelDims(widths=(64, 64, 64, 64, 64), fc_dim=64, vocab_size=64, rnn_size=64, encoding_size=64, seq_length=2)
    c = Context(d, max_image_hw=(64, 64), max_proposals=64)
    yield c
    c.close()


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def rel_err(got, ref):
    scale = ref.abs().max().item() + 1e-30
    return ((got.double() - ref).abs().max().item()) / scale


def run_linear(ctx, M, N, K, relu, seed=0, bias=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / np.sqrt(K)).cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    y = torch.full((M, N), float("nan"), device="cuda")
    ctx.check(ctx._lib.dc_linear(ctx.handle, p(x), p(w), p(b), p(y), M, N, K, relu, None))
    ref = x.double() @ w.double().t()
    if bias:
        ref = ref + b.double()
    if relu:
        ref = ref.clamp_min(0)
    return y, ref


@pytest.mark.parametrize("M,N,K,relu", [(128, 64, 64, 0), (128, 128, 64, 0), (128, 128, 256, 1), (1, 64, 64, 0),
                                        (300, 200, 512, 1), (1000, 72, 256, 0), (257, 1030, 1024, 0), (1000, 4096, 4096, 1)])
def test_tc_linear(ctx, M, N, K, relu):
    y, ref = run_linear(ctx, M, N, K, relu, seed=M + N + K)
    assert not torch.isnan(y).any(), "output rows left unwritten"
    e = rel_err(y, ref)
    assert e < 2e-5, "tcgen05 linear M=%d N=%d K=%d rel err %.3e" % (M, N, K, e)


@pytest.mark.parametrize("M,N,K,relu", [(4000, 4096, 8192, 1), (3900, 4000, 8192, 0)])
def test_tc_linear_tail_split(ctx, M, N, K, relu):
    """256 pair tiles on 74 SM pairs: the 34 tiles of the last round run as a second launch, each cut into two K-ranges, and
    gemm_tail_finalize_kernel adds the ranges (fc6 / fc7 at batch 4).  Ragged M and N in the second case."""
    y, ref = run_linear(ctx, M, N, K, relu, seed=M + N + K)
    assert not torch.isnan(y).any(), "output rows left unwritten"
    e = rel_err(y, ref)
    assert e < 2e-5, "tcgen05 linear M=%d N=%d K=%d rel err %.3e" % (M, N, K, e)


def test_tc_linear_is_fp32_class_accurate(ctx):
    # the split scheme must beat a single fp16 pass by orders of magnitude
    y, ref = run_linear(ctx, 256, 256, 2048, 0, seed=7, bias=False)
    e = rel_err(y, ref)
    assert e < 2e-5, e   # fp32-accumulate class (a single fp16 pass would be ~5e-4)


@pytest.mark.parametrize("M,N,K", [(256, 256, 25088), (300, 128, 25088), (130, 64, 4608)])
def test_tc_linear_long_k_is_fp32_sgemm_class(ctx, M, N, K):
    """The tensor core adds each MMA into its fp32 accumulator with truncation: over fc6's K = 25088 (4704 MMAs per output) a single
    accumulator drifted by 1.2e-4 relative.  With hierarchical accumulation (partial sums of one k-block summed in registers with
    round-to-nearest) the result must be in the class of an fp32 sgemm: a few 1e-7 of the output scale, for every tile shape
    (256-column pair tiles, 128- and 64-column wide tiles)."""
    y, ref = run_linear(ctx, M, N, K, 0, seed=K + N, bias=False)
    e = rel_err(y, ref)
    f32 = rel_err(ref.float(), ref)                                   # rounding the exact result to fp32
    assert e < 2e-6, "tcgen05 linear M=%d N=%d K=%d rel err %.3e (fp32 rounding alone: %.1e)" % (M, N, K, e, f32)


def run_conv(ctx, N, Cin, Cout, H, W, pool, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * np.sqrt(2.0 / (9 * Cin))).cuda()
    b = (torch.randn(Cout, generator=g) * 0.1).cuda()
    oh, ow = ((H + 1) // 2, (W + 1) // 2) if pool else (H, W)
    y = torch.full((N, Cout, oh, ow), float("nan"), device="cuda")
    ctx.check(ctx._lib.dc_conv3x3_relu(ctx.handle, p(x), p(w), p(b), p(y), N, Cin, Cout, H, W, pool, None))
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1).clamp_min(0)
    if pool:
        ref = torch.nn.functional.max_pool2d(ref, 2, 2, 0, ceil_mode=True)
    return y, ref


@pytest.mark.parametrize("N,Cin,Cout,H,W,pool", [
    (1, 64, 64, 8, 16, 0),        # exactly one tile
    (1, 64, 64, 9, 17, 0),        # ragged tile edges
    (1, 64, 128, 38, 45, 0),      # conv5-like map
    (2, 128, 64, 20, 33, 0),      # batch + 2 k-blocks per tap
    (1, 64, 64, 16, 32, 1),       # separate pool kernel
    (1, 64, 64, 16, 32, 2),       # fused pool epilogue, even size
    (1, 64, 128, 75, 90, 2),      # fused ceil-mode pool with odd height (75 -> 38)
    (2, 256, 256, 19, 23, 2),     # batch, odd both, 4 k-blocks per tap
    (1, 64, 64, 150, 200, 2),     # 247 tiles on 148 SMs: resident filter bank reused across tiles of one CTA (conv1_2 path)
    (2, 64, 64, 120, 130, 0),     # same, batch 2, no pool
    (1, 128, 256, 150, 180, 0),   # cta_group::2 pair tiles, several rounds, 2 k-blocks per tap
    (1, 512, 512, 38, 45, 0),     # conv5 shape: 8 channel blocks per tap through the 9-stage B ring
    (1, 256, 256, 144, 160, 0),   # 90 pair tiles on 74 SM pairs: the 16 leftover tiles run as the K-split tail launch + conv_tail_finalize
    (1, 256, 256, 139, 157, 2),   # same with the fused ceil-mode pool (odd sizes) done by the finalize kernel
    (2, 256, 512, 75, 90, 0),     # conv4-like: 120 pair tiles, tail of 46 -> no split (more than half a round), plain path
])
def test_tc_conv(ctx, N, Cin, Cout, H, W, pool):
    y, ref = run_conv(ctx, N, Cin, Cout, H, W, pool, seed=H * W + Cin)
    assert not torch.isnan(y).any(), "output pixels left unwritten"
    e = rel_err(y, ref)
    assert e < 2e-5, "tcgen05 conv rel err %.3e" % e


def test_activation_range_guard(ctx):
    """Split-fp16 planes hold |x| < 65504 only (cvt.satfinite saturates beyond): a layer whose output leaves that range must be
    reported (DC_E_RANGE), not silently clipped; an in-range run on the same context afterwards is clean again."""
    from densecap_b200._lib import DenseCapError
    g = torch.Generator(device="cpu").manual_seed(0)
    x = (torch.rand(1, 64, 16, 16, generator=g) * 4e3).cuda()
    w = torch.ones(64, 64, 3, 3).cuda()
    b = torch.zeros(64).cuda()
    y = torch.empty((1, 64, 16, 16), device="cuda")
    with pytest.raises(DenseCapError, match="fp16 range"):
        ctx.check(ctx._lib.dc_conv3x3_relu(ctx.handle, p(x), p(w), p(b), p(y), 1, 64, 64, 16, 16, 0, None))
    y2, ref = run_conv(ctx, 1, 64, 64, 16, 16, 0, seed=3)
    assert rel_err(y2, ref) < 2e-5
