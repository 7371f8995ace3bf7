"""CPU: the C oracle (the fast checker and the CPU baseline) is bit-identical to the NumPy oracle."""
import numpy as np

from oracle import c_oracle as CO
from oracle import nms_oracle as NO
from oracle import osd_oracle as OO
from oracle import philox_oracle as PO


def test_c_nms_equals_numpy_oracle(code):
    y, cw, _ = PO.gen_frames(3, 0, 300, 2.5, code.G)
    y[5] = 0
    y[6, :7] = 0
    for kw in (dict(), dict(early_stop=True)):
        a = NO.decode(y, code.H, 12, **kw)
        b = CO.nms(y, code.H, 12, float(NO.softplus(-0.048)), traj=True, **kw)
        assert np.array_equal(a["traj"], b["traj"]) and np.array_equal(a["hard"], b["hard"])
        assert np.array_equal(a["iters_used"], b["iters_used"]) and np.array_equal(a["syndrome_nz"], b["syndrome_nz"])
    a = NO.decode(y, code.H, 5, 0.8, 0.9, 1.1)
    b = CO.nms(y, code.H, 5, 0.8, 0.9, 1.1, traj=True)
    assert np.array_equal(a["traj"], b["traj"])


def test_c_osd_equals_numpy_oracle(code):
    y, cw, _ = PO.gen_frames(4, 0, 400, 2.0, code.G)
    fails = np.flatnonzero(CO.nms(y, code.H)["syndrome_nz"])[:12]
    rng = np.random.default_rng(1)
    ys = (y[fails] + rng.normal(0, 0.3, (len(fails), 128))).astype(np.float32)
    yq = y[fails].copy()
    yq[0] = np.round(yq[0] * 4) / 4
    for teps, flags in ((OO.generate_teps_conv(1), 0), (OO.generate_teps_conv(2), 3), (OO.generate_teps_fs(2), 1), (OO.generate_teps_fs(1), 2)):
        r = CO.osd(yq, ys, code.G, OO.pack_teps(teps), flags=flags, truth=cw[fails])
        for n in range(len(fails)):
            o = OO.osd_frame(yq[n], ys[n], code.G, teps, flags=flags, truth=cw[fails[n]])
            assert np.array_equal(o["perm"], r["perm"][n]) and o["best_tep"] == r["best_tep"][n]
            assert o["best_score_q"] == r["best_score_q"][n] and o["score_exp"] == r["score_exp"][n]
            assert np.array_equal(o["codeword"], r["codeword"][n]) and o["truth_score_q"] == r["truth_score_q"][n]
            assert np.array_equal(OO.pack_teps([]) if False else r["redG"][n], np.packbits(o["reduced_G"][:, 64:].astype(np.uint8), axis=1, bitorder="little").view("<u8").reshape(64))
