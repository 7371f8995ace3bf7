"""Scratch timing of the OSD kernels and the pipeline on failure-like frames (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

code = Code(); h = _lib.Handle(code.H, code.G, 0)
A = 0.66943514
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
B = 1 << 21
y = torch.empty((B, 128), dtype=torch.float32, device='cuda'); tr = torch.empty((B, 4), dtype=torch.int32, device='cuda')
h.call('ldpcb_gen_frames', 1, 0, B, 2.5, y, tr, None)
bits = torch.empty((B, 4), dtype=torch.int32, device='cuda'); it = torch.empty(B, dtype=torch.uint8, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda')
t = timeit(lambda: h.call('ldpcb_nms_decode', y, B, 12, A, 1.0, 1.0, 0, bits, it, syn, None, None))
print('nms %.3f ms %.3e f/s' % (t, B / t * 1e3))
fi = syn.bool().nonzero().flatten()
yf = y[fi].contiguous(); Bf = yf.shape[0]
for order in (0, 1, 2):
    t = timeit(lambda: h.call('ldpcb_osd_decode', yf, yf, Bf, order, 0, 0, bits, None, None, None, None, None, None))
    print('osd order %d on %d NMS failures: %.3f ms %.3e f/s' % (order, Bf, t, Bf / t * 1e3))
yr = y[:Bf].contiguous()
for order in (0, 1, 2):
    t = timeit(lambda: h.call('ldpcb_osd_decode', yr, yr, Bf, order, 0, 0, bits, None, None, None, None, None, None))
    print('osd order %d on %d unselected frames: %.3f ms %.3e f/s' % (order, Bf, t, Bf / t * 1e3))
cntr = torch.zeros(16, dtype=torch.int64, device='cuda')
t = timeit(lambda: h.call('ldpcb_decode', y, B, 12, A, 1.0, 1.0, 0, 2, 0, bits, syn, None, tr, cntr, None))
print('decode order 2: %.3f ms %.3e f/s' % (t, B / t * 1e3))
