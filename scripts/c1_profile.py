"""Where a 1000-frame Decoding_model call spends its time (cProfile + wall clock per stage)."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from short_ldpc_decoding_osd_b200 import _lib, globalmap as GL, ms_test
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code
from short_ldpc_decoding_osd_b200 import data_generating as DG
code = Code()
for k, v in dict(code_parameters=code, selected_decoder_type="NMS-1", num_iterations=12).items(): GL.set_map(k, v)
y, lab = DG.testing_data_generating(code, 2.5, 20000, seed=0)
y = np.ascontiguousarray(y, np.float32); lab = np.asarray(lab)
model = ms_test.Decoding_model()
for i in range(3): model(y[:1000], lab[:1000])
for rep in range(3):
    t0 = time.perf_counter()
    for b0 in range(0, 20000, 1000): model(y[b0:b0 + 1000], lab[b0:b0 + 1000])
    print('us per 1000-frame call', (time.perf_counter() - t0) / 20 * 1e6)
pr = cProfile.Profile(); pr.enable()
for b0 in range(0, 20000, 1000): model(y[b0:b0 + 1000], lab[b0:b0 + 1000])
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
