"""Throughput of the DL scheme run entirely on the device (simulate.run_point_dl, BASELINE config 4 shape):
NMS -> trajectories of the detected failures -> DIA FIR -> block minima along the decoding path -> window policy.
Synthetic taps / window classifier (the trained checkpoints are not shipped with the reference)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from short_ldpc_decoding_osd_b200 import _lib, globalmap as GL, nn_net, nn_testing, simulate
from short_ldpc_decoding_osd_b200 import ordered_statistics_decoding as OSD
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

code = Code()
for k, v in dict(code_parameters=code, selected_decoder_type="NMS-1", num_iterations=12, threshold_sum=2, segment_num=6, soft_margin=0.9,
                 decoding_length=30, sliding_win_width=5).items():
    GL.set_map(k, v)
h = _lib.Handle(code.H, code.G, 0)
osd = OSD.osd(code)
path = nn_testing.filter_order_patterns(nn_testing.convention_segment_path())
tep_info = nn_testing.generate_teps(osd, path)
rng = np.random.default_rng(3)
taps = (np.full(13, 1 / 13) + 0.03 * rng.normal(size=13)).astype(np.float32)
net = nn_net.Predict_outlier_light(5, W1=np.eye(6, dtype=np.float32), W2=np.array([[0, -0.5], [0, 0.5], [0, 0], [0, 0], [0, 0], [0, 0.15]], np.float32))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
ebn0 = float(sys.argv[2]) if len(sys.argv) > 2 else 2.5
simulate.run_point_dl(h, ebn0, 1 << 18, tep_info, taps, 0.05, net.W1, net.W2, seed=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
t, out = simulate.run_point_dl(h, ebn0, n, tep_info, taps, 0.05, net.W1, net.W2, seed=2)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"frames": n, "ebn0_db": ebn0, "seconds": dt, "frames_per_s": n / dt, "path_blocks": len(tep_info[0]), "teps_on_path": int(tep_info[1][-1]),
                  "fer_nms": t.fer_nms, "nms_detected": t.nms_detected, **out}))
