"""BASELINE config 1 through the drop-in API: CCSDS (128,64), NMS 12 iterations, Eb/N0 2.5 dB, 10^5 frames, decoded
with ms_test.Decoding_model in batches of 1000 as ldpc_128_testing.py:20,117-131 does, and in one call of 10^5.
Host NumPy arrays in, (FER, BER, undetected, 13-row retest buffer of the detected failures) out, like the reference."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from short_ldpc_decoding_osd_b200 import _lib, globalmap as GL, ms_test
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
from short_ldpc_decoding_osd_b200 import data_generating as DG

code = Code()
for k, v in dict(code_parameters=code, selected_decoder_type="NMS-1", num_iterations=12).items():
    GL.set_map(k, v)
F = 100000
y, lab = DG.testing_data_generating(code, 2.5, F, seed=0)
y = np.ascontiguousarray(y, np.float32); lab = np.asarray(lab)
model = ms_test.Decoding_model()
model(y[:1000], lab[:1000])  # warm-up: library load, handle creation
out = {}
for name, bs in (("batches_of_1000", 1000), ("one_call", F)):
    t0 = time.perf_counter()
    fer = und = 0.0; rows = 0
    for b0 in range(0, F, bs):
        f, ber, u, buf = model(y[b0:b0 + bs], lab[b0:b0 + bs])
        fer += f * min(bs, F - b0); und += u; rows += len(buf[0])
    dt = time.perf_counter() - t0
    out[name] = {"seconds": dt, "frames_per_s": F / dt, "fer": fer / F, "undetected": int(und), "retest_rows": rows}
print(json.dumps({"config": "C1: NMS-1 12 it, 2.5 dB, 1e5 frames via ms_test.Decoding_model (host arrays in/out)", **out}))
