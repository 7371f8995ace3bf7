"""Timing of ldpcb_osd_block_minima (DL path, convention path with sum w <= 2: 27 blocks, 2081 TEPs) on NMS failures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from short_ldpc_decoding_osd_b200 import _lib, globalmap as GL, nn_testing
from short_ldpc_decoding_osd_b200 import ordered_statistics_decoding as OSD
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
code = Code()
for k, v in dict(code_parameters=code, selected_decoder_type="NMS-1", num_iterations=12, threshold_sum=2, segment_num=6, soft_margin=0.9, decoding_length=30, sliding_win_width=5).items():
    GL.set_map(k, v)
h = _lib.Handle(code.H, code.G, 0)
teps_list, acc = nn_testing.generate_teps(OSD.osd(code), nn_testing.filter_order_patterns(nn_testing.convention_segment_path()))
packed = torch.from_numpy(np.concatenate([OSD.pack_dl_teps(b) for b in teps_list]).view(np.int32)).cuda()
starts = torch.from_numpy(np.asarray(acc, dtype=np.int32)).cuda(); nb = len(teps_list)
B = 1 << 20; A = 0.66943514
y = torch.empty((B, 128), dtype=torch.float32, device='cuda'); tr = torch.empty((B, 4), dtype=torch.int32, device='cuda')
h.call('ldpcb_gen_frames', 2, 0, B, 2.5, y, tr, None)
bits = torch.empty((B, 4), dtype=torch.int32, device='cuda'); syn = torch.empty(B, dtype=torch.uint8, device='cuda'); met = torch.empty((B, 128), dtype=torch.float32, device='cuda')
taps = (np.full(13, 1 / 13)).astype(np.float32)
h.call('ldpcb_nms_decode_fir', y, B, 12, A, 1.0, 1.0, taps, 0.0, bits, syn, met, None)
m = syn.bool(); yf = y[m].contiguous(); mf = met[m].contiguous(); tf = tr[m].contiguous(); n = yf.shape[0]
bm = torch.empty((n, nb), dtype=torch.int64, device='cuda'); ex = torch.empty(n, dtype=torch.int32, device='cuda'); ts = torch.empty(n, dtype=torch.int64, device='cuda')
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); t = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record(); fn(); b.record(); torch.cuda.synchronize(); t.append(a.elapsed_time(b))
    return min(t)
for hint in (0, 2):
    fl = OSD.FLAGS_DL | (hint << _lib.OSD_MAXW_SHIFT)
    t = timeit(lambda: h.call('ldpcb_osd_block_minima', mf, yf, n, packed, int(packed.numel()), starts, nb, fl, bm, None, ex, tf, ts, None, None))
    print(f'block minima, max-weight hint {hint}: {n} frames, {nb} blocks: {t:.3f} ms  ({t * 230000 / n:.3f} ms per 2.3e5)  {n / t * 1e3:.3e} frames/s')
t = timeit(lambda: h.call('ldpcb_osd_decode', yf, yf, n, 2, 0, 0, bits, None, None, None, None, None, None))
print(f'for scale: exhaustive order 2 on the same frames {t:.3f} ms')
