"""Order-3 OSD on NMS failures: tensor-core sweep (osd3.cu) vs the generic kernel, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
code = Code(); h = _lib.Handle(code.H, code.G, 0)
A = 0.66943514
def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
y = torch.empty((B, 128), dtype=torch.float32, device='cuda'); tr = torch.empty((B, 4), dtype=torch.int32, device='cuda')
h.call('ldpcb_gen_frames', 2024, 0, B, 2.5, y, tr, None)
bits = torch.empty((B, 4), dtype=torch.int32, device='cuda'); it = torch.empty(B, dtype=torch.uint8, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda')
h.call('ldpcb_nms_decode', y, B, 12, A, 1.0, 1.0, 0, bits, it, syn, None, None)
yf = y[syn.bool()].contiguous(); n = yf.shape[0]
for tag in ('tensor', 'generic'):
    if tag == 'generic': os.environ['LDPCB_OSD3_GENERIC'] = '1'
    m = n if tag == 'tensor' else n // 4
    t = timeit(lambda: h.call('ldpcb_osd_decode', yf, yf, m, 3, 0, 0, bits, None, None, None, None, None, None))
    print(f'order 3 {tag}: {m} frames {t:.3f} ms  {m / t * 1e3:.3e} frames/s')
os.environ.pop('LDPCB_OSD3_GENERIC', None)
if len(sys.argv) > 2:  # single launch for ncu
    h.call('ldpcb_osd_decode', yf, yf, min(n, 32768), 3, 0, 0, bits, None, None, None, None, None, None); torch.cuda.synchronize()
print('done')
