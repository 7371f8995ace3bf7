"""A/B of the PB-OSD kernel's register budget / occupancy (env LDPCB_PB_MINB, LDPCB_PB_NOCAP3), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
code = Code(); A = 0.66943514
def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
B = 1 << 20
for ebn0 in (3.0, 2.5):
    h = _lib.Handle(code.H, code.G, 0)
    y = torch.empty((B, 128), dtype=torch.float32, device='cuda'); tr = torch.empty((B, 4), dtype=torch.int32, device='cuda')
    h.call('ldpcb_gen_frames', 2025, 0, B, ebn0, y, tr, None)
    bits = torch.empty((B, 4), dtype=torch.int32, device='cuda'); it = torch.empty(B, dtype=torch.uint8, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda')
    h.call('ldpcb_nms_decode', y, B, 12, A, 1.0, 1.0, 0, bits, it, syn, None, None)
    yf = y[syn.bool()].contiguous(); n = yf.shape[0]
    st4 = torch.empty((n, 4), dtype=torch.int32, device='cuda')
    for minb in ('6', '8', '10'):
        for nocap in (False, True):
            os.environ['LDPCB_PB_MINB'] = minb
            if nocap: os.environ['LDPCB_PB_NOCAP3'] = '1'
            else: os.environ.pop('LDPCB_PB_NOCAP3', None)
            for order, m in ((2, n), (3, min(n, 65536))):
                if order == 2 and nocap: continue
                t = timeit(lambda: h.call('ldpcb_osd_pb_decode', yf, m, order, ebn0, bits, st4, None, None, None))
                print(f'{ebn0} dB minb {minb} nocap3 {int(nocap)} order {order}: {m} frames {t:.3f} ms  {m / t * 1e3:.3e} frames/s', flush=True)
    h.close()
