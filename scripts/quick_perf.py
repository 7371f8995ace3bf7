"""Scratch timing of the individual kernels (CUDA events on the launching stream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

code = Code(); h = _lib.Handle(code.H, code.G, 0)
ALPHA = 0.66943514
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts)//2]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
y = torch.empty((B,128), dtype=torch.float32, device='cuda'); tr = torch.empty((B,4), dtype=torch.int32, device='cuda')
print('gen', B, timeit(lambda: h.call('ldpcb_gen_frames', 1, 0, B, 2.5, y, tr, None)))
bits = torch.empty((B,4), dtype=torch.int32, device='cuda'); it = torch.empty(B, dtype=torch.uint8, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda')
for early in (0,1):
    t = timeit(lambda: h.call('ldpcb_nms_decode', y, B, 12, ALPHA, 1.0, 1.0, early, bits, it, syn, None, None))
    print('nms early=%d'%early, t, 'frames/s %.3e' % (B / t[0] * 1e3))
fails = syn.bool().sum().item(); print('fail frac', fails / B)
idx = torch.empty(B, dtype=torch.int32, device='cuda'); cnt = torch.empty(1, dtype=torch.int32, device='cuda')
print('select', timeit(lambda: h.call('ldpcb_select_flagged', syn, B, idx, cnt, None)))
Bo = min(B, 1 << 18)
yo = y[:Bo]
for order in (0, 1, 2, 3):
    Bx = Bo if order < 3 else Bo // 16
    t = timeit(lambda: h.call('ldpcb_osd_decode', yo, yo, Bx, order, 0, 0, bits, None, None, None, None, None, None), n=3, warm=1)
    print('osd order', order, t, 'frames/s %.3e' % (Bx / t[0] * 1e3))
cntr = torch.zeros(16, dtype=torch.int64, device='cuda')
for order in (-1, 1, 2):
    t = timeit(lambda: h.call('ldpcb_decode', y, B, 12, ALPHA, 1.0, 1.0, 0, order, 0, bits, syn, None, tr, cntr, None), n=3, warm=1)
    print('decode order', order, t, 'frames/s %.3e' % (B / t[0] * 1e3))
t = timeit(lambda: h.call('ldpcb_simulate', 1, 0, B, 2.5, 12, ALPHA, 1.0, 1.0, 0, 2, 0, cntr, None), n=3, warm=1)
print('simulate order 2', t, 'frames/s %.3e' % (B / t[0] * 1e3))
nt = torch.empty(Bo, dtype=torch.int32, device='cuda'); sk = torch.empty(Bo, dtype=torch.uint8, device='cuda')
fails_idx = syn.bool().nonzero().flatten()[:Bo]
yf = y[fails_idx].contiguous(); Bf = yf.shape[0]
for order in (1, 2, 3):
    t = timeit(lambda: h.call('ldpcb_osd_fs_decode', yf, Bf, order, 6.5, 30, 6.4, bits, None, nt, sk, None, None, None, None), n=3, warm=1)
    print('fs-osd order', order, t, 'frames/s %.3e' % (Bf / t[0] * 1e3), 'avg teps %.1f' % nt[:Bf].float().mean().item())
st4 = torch.empty((Bf, 4), dtype=torch.int32, device='cuda')
for order in (1, 2, 3):
    Bp = Bf if order == 1 else Bf // 4
    t = timeit(lambda: h.call('ldpcb_osd_pb_decode', yf, Bp, order, 2.5, bits, st4, None, None, None), n=3, warm=1)
    print('pb-osd order', order, t, 'frames/s %.3e' % (Bp / t[0] * 1e3), 'avg teps %.1f' % st4[:Bp, 0].float().mean().item())
