"""One launch of each hot kernel at a fixed size, for ncu (python scripts/profile_case.py [frames]).  Writes the
build stamp of the library it ran to gpurun_out/prof_stamp.json (scripts/make_traffic.py needs it)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from short_ldpc_decoding_osd_b200 import _lib, build
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
code = Code(); h = _lib.Handle(code.H, code.G, 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
A = 0.66943514
y = torch.empty((B, 128), dtype=torch.float32, device='cuda'); tr = torch.empty((B, 4), dtype=torch.int32, device='cuda')
h.call('ldpcb_gen_frames', 1, 0, B, 2.5, y, tr, None)
bits = torch.empty((B, 4), dtype=torch.int32, device='cuda'); it = torch.empty(B, dtype=torch.uint8, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda')
B3 = B // 8
for rep in range(2):
    h.call('ldpcb_nms_decode', y, B, 12, A, 1.0, 1.0, 0, bits, it, syn, None, None)
    for order in (1, 2):
        h.call('ldpcb_osd_decode', y, y, B, order, 0, 0, bits, None, None, None, None, None, None)
    h.call('ldpcb_osd_decode', y, y, B3, 3, 0, 0, bits, None, None, None, None, None, None)
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump({"build_stamp": build._stamp(), "frames": B, "frames_of": {"osd_kernel_order3": B3}}, open(os.path.join(ROOT, 'gpurun_out', 'prof_stamp.json'), 'w'))
print('done'); h.close()
