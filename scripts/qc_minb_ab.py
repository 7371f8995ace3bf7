"""A/B of the nms_qc register budget (2 vs 3 CTAs per SM) and of the whole pipeline, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
code = Code(); A = 0.66943514
def timeit(fn, n=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
B = 1 << 21
hs = {}
for m in ("3", "2"):
    os.environ["LDPCB_QC_MINB"] = m; hs[m] = _lib.Handle(code.H, code.G, 0)
del os.environ["LDPCB_QC_MINB"]
y = torch.empty((B, 128), dtype=torch.float32, device='cuda'); tr = torch.empty((B, 4), dtype=torch.int32, device='cuda')
hs["3"].call('ldpcb_gen_frames', 2024, 0, B, 2.5, y, tr, None)
bits = torch.empty((B, 4), dtype=torch.int32, device='cuda'); it = torch.empty(B, dtype=torch.uint8, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda')
cnt = torch.zeros(16, dtype=torch.int64, device='cuda')
for m, h in hs.items():
    t = timeit(lambda: h.call('ldpcb_nms_decode', y, B, 12, A, 1.0, 1.0, 0, bits, it, syn, None, None))
    print(f'minb {m} nms: {t:.3f} ms  {B / t * 1e3:.3e} frames/s')
    t = timeit(lambda: h.call('ldpcb_decode', y, B, 12, A, 1.0, 1.0, 0, 2, 0, bits, syn, None, tr, cnt, None))
    print(f'minb {m} decode order 2: {t:.3f} ms  {B / t * 1e3:.3e} frames/s')
    t = timeit(lambda: h.call('ldpcb_simulate', 7, 0, B, 2.5, 12, A, 1.0, 1.0, 0, 2, 0, cnt, None))
    print(f'minb {m} simulate order 2: {t:.3f} ms  {B / t * 1e3:.3e} frames/s')
    for Bs in (1024, 4096):
        t = timeit(lambda: h.call('ldpcb_decode', y, Bs, 12, A, 1.0, 1.0, 0, 2, 0, bits, syn, None, tr, cnt, None), n=20, warm=5)
        print(f'minb {m} decode order 2, {Bs} frames: {t * 1e3:.1f} us')
