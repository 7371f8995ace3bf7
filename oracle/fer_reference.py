"""FER reference points (TEST INFRASTRUCTURE): frames from the REFERENCE'S OWN generator
(/root/reference/LDPC_128/Testing_data_gen_128/data_generating.py, NumPy MT19937, seeded) decoded by the C
oracle (bit-identical to the NumPy restatement of the reference graph, which is bit-identical to the
reference source under the TF shim).  Writes tests/golden/fer_reference.json with Wilson 95% intervals.
Only runs in the build container (needs /root/reference)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/LDPC_128/Testing_data_gen_128")

import contextlib  # noqa: E402
import io  # noqa: E402

import data_generating as RefGen  # noqa: E402  (the reference module)
import fill_matrix_info as RefFill  # noqa: E402
import globalmap as RefGL  # noqa: E402

from oracle import c_oracle as CO  # noqa: E402
from oracle import osd_oracle as OO  # noqa: E402
from short_ldpc_decoding_osd_b200.simulate import wilson_interval  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    code = RefFill.Code("/root/reference/LDPC_128/Testing_data_gen_128/CCSDS_ldpc_n128_k64.alist")
RefGL.set_map("Rayleigh_fading", False)
RefGL.set_map("ALL_ZEROS_CODEWORD_TESTING", False)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
out = {"frames_per_point": N, "alpha": 0.66943514, "iters": 12, "generator": "reference testing_data_generating, np.random.seed(2000+i)", "points": []}
teps = {o: OO.pack_teps(OO.generate_teps_conv(o)) for o in (1, 2)}
for i, snr in enumerate([2.0, 2.5, 3.0, 3.5]):
    np.random.seed(2000 + i)
    y, lab = RefGen.testing_data_generating(code, snr, N)
    y = y.astype(np.float32)
    r = CO.nms(y, code.H, 12, 0.66943514)
    err = (r["hard"] != lab).any(axis=1)
    fails = np.flatnonzero(r["syndrome_nz"])
    pt = {"ebn0_db": snr, "frames": N, "nms_frame_err": int(err.sum()), "nms_detected": int(len(fails)),
          "nms_undetected": int((err & ~r["syndrome_nz"]).sum())}
    pt["fer_nms"] = pt["nms_frame_err"] / N
    pt["fer_nms_ci95"] = wilson_interval(pt["nms_frame_err"], N)
    for o in (1, 2):
        d = CO.osd(np.ascontiguousarray(y[fails]), None, code.G, teps[o], want_perm=False)
        osd_err = int((d["codeword"] != lab[fails]).any(axis=1).sum())
        fin = osd_err + pt["nms_undetected"]
        pt[f"final_frame_err_osd{o}"] = fin
        pt[f"fer_final_osd{o}"] = fin / N
        pt[f"fer_final_osd{o}_ci95"] = wilson_interval(fin, N)
    out["points"].append(pt)
    print(pt, flush=True)
with open(os.path.join(ROOT, "tests", "golden", "fer_reference.json"), "w") as f:
    json.dump(out, f, indent=1)
