"""One launch of the PB-OSD kernel on NMS failures, for ncu (python scripts/profile_pb.py [order] [frames])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
code = Code(); h = _lib.Handle(code.H, code.G, 0)
order = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 1 << 19
y = torch.empty((B, 128), dtype=torch.float32, device='cuda'); tr = torch.empty((B, 4), dtype=torch.int32, device='cuda')
h.call('ldpcb_gen_frames', 1, 0, B, 2.5, y, tr, None)
bits = torch.empty((B, 4), dtype=torch.int32, device='cuda'); it = torch.empty(B, dtype=torch.uint8, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda')
h.call('ldpcb_nms_decode', y, B, 12, 0.66943514, 1.0, 1.0, 0, bits, it, syn, None, None)
yf = y[syn.bool()].contiguous(); Bf = min(yf.shape[0], int(sys.argv[2]) if len(sys.argv) > 2 else 32768)
st = torch.empty((Bf, 4), dtype=torch.int32, device='cuda')
for _ in range(2):
    h.call('ldpcb_osd_pb_decode', yf, Bf, order, 2.5, bits, st, None, None, None)
torch.cuda.synchronize(); print('done', Bf, st[:, 0].float().mean().item()); h.close()
