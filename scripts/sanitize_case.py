"""Smallest run that touches every kernel once (for compute-sanitizer; never timed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

code = Code(); h = _lib.Handle(code.H, code.G, 0)
B = 259; A = 0.66943514
y = torch.empty((B, 128), dtype=torch.float32, device='cuda'); tr = torch.empty((B, 4), dtype=torch.int32, device='cuda')
h.call('ldpcb_gen_frames', 1, 0, B, 2.0, y, tr, None)
bits = torch.empty((B, 4), dtype=torch.int32, device='cuda'); it = torch.empty(B, dtype=torch.uint8, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda')
traj = torch.empty((B, 13, 128), dtype=torch.float32, device='cuda')
for early in (0, 1):
    h.call('ldpcb_nms_decode', y, B, 12, A, 1.0, 1.0, early, bits, it, syn, traj, None)
    h.call('ldpcb_nms_decode', y, B, 12, A, 0.9, 1.1, early, bits, it, syn, None, None)
bt = torch.empty(B, dtype=torch.int32, device='cuda'); bq = torch.empty(B, dtype=torch.int64, device='cuda'); ex = torch.empty(B, dtype=torch.int32, device='cuda')
pm = torch.empty((B, 128), dtype=torch.uint8, device='cuda'); rg = torch.empty((B, 64), dtype=torch.int64, device='cuda')
yq = torch.round(y * 4) / 4
for order in (0, 1, 2):
    for src in (y, yq):
        h.call('ldpcb_osd_decode', src, src, B, order, order & 1, order, bits, bt, bq, ex, pm, rg, None)
for src, n3, to in ((y, 259, 0), (yq, 33, 1), (y, 3, 1)):  # order 3: tensor-core sweep, its exact fallback (quantised input), both TEP orders
    h.call('ldpcb_osd_decode', src, src, n3, 3, to, 3 * to, bits, bt, bq, ex, pm, rg, None)
os.environ['LDPCB_OSD3_GENERIC'] = '1'
h.call('ldpcb_osd_decode', y, y, 17, 3, 0, 0, bits, bt, bq, ex, pm, rg, None)
del os.environ['LDPCB_OSD3_GENERIC']
nt = torch.empty(B, dtype=torch.int32, device='cuda'); sk = torch.empty(B, dtype=torch.uint8, device='cuda')
h.call('ldpcb_osd_fs_decode', y, B, 2, 6.5, 30, 6.4, bits, bt, nt, sk, bq, ex, pm, None)
h.call('ldpcb_osd_fs_decode', y, B, 3, 6.5, 30, 6.4, bits, bt, nt, sk, bq, ex, pm, None)   # three launches: defer, tensor class 3, exact rest
h.call('ldpcb_osd_fs_decode', yq, 77, 3, 9.0, 26, 3.2, bits, bt, nt, sk, None, None, None, None)
st4 = torch.empty((B, 4), dtype=torch.int32, device='cuda')
for order in (1, 2, 3):
    h.call('ldpcb_osd_pb_decode', y, 61 if order == 3 else B, order, 2.5, bits, st4, bq, ex, None)
teps = h.tep_table(2, 0); starts = np.array([0, 1, 65, 700, 2081], dtype=np.int32)
bm = torch.empty((B, 4), dtype=torch.int64, device='cuda'); ba = torch.empty((B, 4), dtype=torch.int32, device='cuda'); ts = torch.empty(B, dtype=torch.int64, device='cuda')
h.call('ldpcb_osd_block_minima', yq, y, B, torch.from_numpy(teps.view(np.int32)).cuda(), len(teps), torch.from_numpy(starts).cuda(), 4, 3, bm, ba, ex, tr, ts, pm, None)
cnt = torch.zeros(16, dtype=torch.int64, device='cuda')
for order in (-1, 0, 1, 2, 3):  # fused two-kernel pipeline (fixed iterations) and the seven-launch one (early stop)
    for early in (0, 1):
        h.call('ldpcb_decode', y, B if order < 3 else 41, 12, A, 1.0, 1.0, early, order, 0, bits, syn, bt, tr, cnt, None)
h.call('ldpcb_decode', y, 1, 12, A, 1.0, 1.0, 0, 2, 0, bits, None, None, None, None, None)
for early, order, n in ((1, 2, 1000), (0, 2, 1001), (0, -1, 77), (0, 3, 64), (0, 1, 1)):
    h.call('ldpcb_simulate', 3, 7, n, 2.0, 12, A, 1.0, 1.0, early, order, 0, cnt, None)
met = torch.empty((B, 128), dtype=torch.float32, device='cuda')
h.call('ldpcb_nms_decode_fir', y, B, 12, A, 1.0, 1.0, np.linspace(0.1, 0.2, 13).astype(np.float32), 0.05, bits, syn, met, None)
out = torch.empty((B, 128), dtype=torch.float32, device='cuda')
h.call('ldpcb_dia_fir', traj, B, 13, np.ones(13, np.float32) / 13, 0.1, out, None)
idx = torch.empty(B, dtype=torch.int32, device='cuda'); c1 = torch.empty(1, dtype=torch.int32, device='cuda')
h.call('ldpcb_select_flagged', syn, B, idx, c1, None); h.call('ldpcb_gather_rows', y, idx, c1, B, 128, out, None)
torch.cuda.synchronize()
yh = y.cpu().numpy(); bh = np.empty((B, 4), np.uint32); sh = np.empty(B, np.uint8); th = tr.cpu().numpy().view(np.uint32); ch = np.zeros(16, np.uint64); bth = np.empty(B, np.int32)
h.call('ldpcb_decode_host', yh, B, 12, A, 1.0, 1.0, 0, 2, 0, bh, sh, bth, th, ch)
h.call('ldpcb_osd_fs_decode_host', yh, B, 1, 6.5, 30, 6.4, bh, bth, None, None)
fi = np.empty(50, np.int32); ft = np.empty((50, 13, 128), np.float32); nf = np.zeros(1, np.int64)
h.call('ldpcb_nms_retest_host', yh, B, 12, A, 1.0, 1.0, th, bh, sh, ch, 50, fi, ft, nf)
print('sanitize case done', ch[:12], int(c1.item()))
h.close()
