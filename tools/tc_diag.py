"""Diagnostic for the tcgen05 GEMM: structured inputs that reveal layout/descriptor mistakes."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from densecap_b200.runtime import Context
from densecap_b200.synth import ModelDims
d = ModelDims(widths=(64,)*5, fc_dim=64, vocab_size=64, rnn_size=64, encoding_size=64, seq_length=2)
ctx = Context(d, max_image_hw=(64, 64), max_proposals=64, engine=1)
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
def run(M, N, K, x, w, eng):
    y = torch.full((M, N), float('nan'), device='cuda')
    rc = ctx._lib.dc_linear(ctx.handle, p(x), p(w), None, p(y), M, N, K, 0, eng, None)
    if rc: print('rc', rc, ctx._lib.dc_last_error(ctx.handle))
    torch.cuda.synchronize(); return y
for (M, N, K) in [(128, 64, 64), (128, 128, 64), (128, 128, 128), (256, 256, 192)]:
    print('--- M,N,K', M, N, K)
    # 1) A = one-hot rows (row m has 1 at k = m % K), W[n,k] = n*1000 + k  -> y[m,n] = n*1000 + (m%K)
    x = torch.zeros(M, K, device='cuda'); x[torch.arange(M), torch.arange(M) % K] = 1
    w = (torch.arange(N, device='cuda')[:, None] * 1000 + torch.arange(K, device='cuda')[None, :]).float()
    ref = x.double() @ w.double().t()
    for eng, nm in [(1, 'simt'), (0, 'tc')]:
        y = run(M, N, K, x, w, eng)
        bad = (y.double() - ref).abs() > 0.5
        print(nm, 'onehot: nan', int(torch.isnan(y).sum()), 'bad', int(bad.sum()), 'of', y.numel())
        if bad.any() and nm == 'tc':
            idx = bad.nonzero()[:8].tolist()
            for (m, n) in idx: print('   y[%d,%d]=%s expected %s' % (m, n, y[m, n].item(), ref[m, n].item()))
            print('   row0[:8]', y[0, :8].tolist()); print('   col0[:8]', y[:8, 0].tolist())
    # 2) random
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda')
    ref = x.double() @ w.double().t()
    for eng, nm in [(1, 'simt'), (0, 'tc')]:
        y = run(M, N, K, x, w, eng)
        print(nm, 'random: rel err %.3e' % ((y.double() - ref).abs().max() / ref.abs().max()).item())
ctx.close()
