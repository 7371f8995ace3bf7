"""Golden-vector generator (TEST INFRASTRUCTURE): runs the REFERENCE'S OWN source files, unmodified, from
/root/reference under the NumPy TensorFlow emulation in oracle/tf_shim/, on seeded inputs, and writes
tests/golden/*.npz.  /root/reference only exists in the build container, so the fixtures are committed
and the tests read the fixtures.

    python oracle/ref_runner.py all          # regenerate every fixture (a few minutes)
    python oracle/ref_runner.py nms|osd|fs|pb|dl|gf2|gen
    python oracle/ref_runner.py nms --real-tf --out /tmp/g   # same run under a REAL TensorFlow (no shim on sys.path),
                                                             # fixture written to /tmp/g: tests/test_tf_ready.py diffs
                                                             # it against the committed shim fixture
The reference tree is /root/reference, or $LDPCB_REFERENCE_ROOT.

Each sub-command runs in its own process with sys.path = [oracle/tf_shim, <one reference directory>]
because the reference's directories all define modules with the same names (globalmap, fill_matrix_info,
convention_osd, ...).
"""
from __future__ import annotations

import ast
import contextlib
import io
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(os.environ.get("LDPCB_REFERENCE_ROOT", "/root/reference"), "LDPC_128")
GOLD = os.path.join(ROOT, "tests", "golden")
REAL_TF = False
ALIST = "CCSDS_ldpc_n128_k64.alist"


def _enter(refdir: str):
    sys.path.insert(0, os.path.join(REF, refdir))
    if not REAL_TF:
        sys.path.insert(0, os.path.join(HERE, "tf_shim"))
    os.chdir(tempfile.mkdtemp(prefix="refrun_"))  # the reference writes ./log/*.txt relative to cwd


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _frames(code, snr, n, seed):
    """The reference's own generator (NumPy only), seeded; stored as float32 like its TFRecords."""
    sys.path.insert(0, os.path.join(REF, "Testing_data_gen_128"))
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_gen", os.path.join(REF, "Testing_data_gen_128", "data_generating.py"))
    mod = importlib.util.module_from_spec(spec)
    import globalmap as GL

    GL.set_map("Rayleigh_fading", False)
    GL.set_map("ALL_ZEROS_CODEWORD_TESTING", False)
    spec.loader.exec_module(mod)
    np.random.seed(seed)
    y, lab = mod.testing_data_generating(code, snr, n)
    return y.astype(np.float32), lab.astype(np.int64)


def run_gen():
    """Reference frame generator + Code class -> H, G and a seeded batch (inputs for the other fixtures)."""
    _enter("Testing_data_gen_128")
    import fill_matrix_info as F

    with _quiet():
        code = F.Code(os.path.join(REF, "Testing_data_gen_128", ALIST))
    y, lab = _frames(code, 2.5, 4000, 0)
    np.savez_compressed(os.path.join(GOLD, "code_ref.npz"), H=code.H.astype(np.uint8), G=code.G.astype(np.uint8), k=code.k,
                        max_chk_degree=code.max_chk_degree)
    # moments of the reference generator for the statistical comparison with the Philox kernel
    z = (np.where(lab == 0, y, -y) - 1.0)
    np.savez_compressed(os.path.join(GOLD, "gen_ref_stats.npz"), snr=2.5, n=y.shape[0], sigma_hat=z.std(), mean_hat=z.mean(),
                        ones_frac=lab.mean(), y_head=y[:8], labels_head=lab[:8].astype(np.uint8))


def run_nms():
    _enter("Ldpc_128_testing")
    import fill_matrix_info as F
    import globalmap as GL

    with _quiet():
        code = F.Code(os.path.join(REF, "Ldpc_128_testing", ALIST))
    GL.set_map("selected_decoder_type", "NMS-1")
    GL.set_map("num_iterations", 12)
    GL.set_map("code_parameters", code)
    import ms_test as R  # the reference module, unmodified

    y, lab = _frames(code, 2.5, 96, 1)
    # a few hand-made rows: exact zeros (tf.sign(0)=0), noiseless, huge and tiny magnitudes
    y[90] = 0.0
    y[91, :9] = 0.0
    y[92] = np.where(lab[92] == 0, 1.0, -1.0)
    y[93] *= 1e20
    y[94] *= 1e-30
    model = R.Decoding_model()
    with _quiet():
        fer, ber, undetected, buffer = model(y, lab)
        soft = model.layer(y, lab)
        fer2, ber2, und2, index = model.get_eval(soft, lab)
    alpha = np.asarray(__import__("tensorflow").nn.softplus(model.layer.shared_check_weight))
    np.savez_compressed(
        os.path.join(GOLD, "nms_ref_shim.npz"), y=y, labels=lab.astype(np.uint8), soft=np.stack([np.asarray(s) for s in soft], 1),
        fer=float(fer), ber=float(np.asarray(ber)), undetected=int(undetected), index=np.asarray(index).astype(np.int64),
        n_buffer_rows=len(buffer[0]), buffer_first=np.stack([np.asarray(b) for b in buffer[0][:13]]), alpha=alpha,
        raw_weight=np.asarray(model.layer.shared_check_weight))


def _failed(code, n, seed, snr=2.5):
    """Frames on which the reference NMS (under the shim) leaves a non-zero syndrome: the OSD inputs."""
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    from oracle import nms_oracle as NO

    y, lab = _frames(code, snr, 8 * n, seed)
    r = NO.decode(y, code.H, 12)
    idx = np.flatnonzero(r["syndrome_nz"])[:n]
    return y[idx], lab[idx]


def run_osd():
    """PB_OSD: swapped_info / identify_mrb / full_gf2elim; FS_OSD copy of convention_osd_main."""
    _enter("FS_OSD")
    import fill_matrix_info as F
    import globalmap as GL
    import tensorflow as tf

    with _quiet():
        code = F.Code(os.path.join(REF, "FS_OSD", ALIST))
    GL.set_map("code_parameters", code)
    GL.set_map("order_limit", 2)
    import convention_osd as C
    import fs_testing as R  # swapped_info is identical in pb_testing.py and fs_testing.py

    y, lab = _failed(code, 40, 2)
    y[36] = np.round(y[36] * 4) / 4  # ties in |y|
    y[37, ::5] = 0.0
    captured = {}
    real_argmin = tf.argmin

    def spy(x, *a, **k):
        captured["scores"] = np.asarray(x).copy()
        return real_argmin(x, *a, **k)

    tf.argmin = spy
    out = {k: [] for k in ("upd_in", "upd_lab", "redG", "ok1", "ok2", "phase1", "phase2", "idx1", "idx2", "gap1", "gap2")}
    teps = {o: np.asarray(C.generate_teps(o)) for o in (0, 1, 2)}
    bnd = {o: C.query_boundary(o) for o in (0, 1, 2)}
    for i in range(len(y)):
        ui, ul, rg = R.swapped_info(y[i], lab[i])
        out["upd_in"].append(np.asarray(ui))
        out["upd_lab"].append(np.asarray(ul).astype(np.uint8))
        out["redG"].append(np.packbits(np.asarray(rg).astype(np.uint8), axis=1))
        for o in (1, 2):
            ok, T, phase = C.convention_osd_main((ui, ul, rg, teps[o], bnd[o]))
            s = np.sort(captured["scores"])
            out[f"ok{o}"].append(bool(ok))
            out[f"phase{o}"].append(int(phase))
            out[f"idx{o}"].append(int(np.argmin(captured["scores"])))
            out[f"gap{o}"].append(float(s[1] - s[0]))
    np.savez_compressed(
        os.path.join(GOLD, "osd_ref_shim.npz"), y=y, labels=lab.astype(np.uint8),
        teps0=np.packbits(teps[0].astype(np.uint8), axis=1), teps1=np.packbits(teps[1].astype(np.uint8), axis=1),
        teps2=np.packbits(teps[2].astype(np.uint8), axis=1), boundary2=np.array(bnd[2]),
        **{k: np.array(v) for k, v in out.items()})


def run_fs():
    """FS_OSD: TEP order, per-frame fs_osd outcome (S/F and number of TEPs) parsed from its own log."""
    _enter("FS_OSD")
    import fill_matrix_info as F
    import globalmap as GL

    with _quiet():
        code = F.Code(os.path.join(REF, "FS_OSD", ALIST))
    GL.set_map("code_parameters", code)
    GL.set_map("termination_num_threshlod", 100)
    GL.set_map("convention_osd", False)
    GL.set_map("miracle_view", False)
    GL.set_map("fs_osd", True)
    GL.set_map("d_min", 14)
    GL.set_map("tau_psc", 30)
    import fs_testing as R

    class DS:
        def __init__(self, y, lab):
            self.b = [(y[None, :], lab[None, :])]

        def as_numpy_iterator(self):
            return iter(self.b)

    y, lab = _failed(code, 30, 3)
    res = {}
    for order in (1, 2, 3):  # order_limit 3 is the reference's default (FS_OSD/globalmap.py:44): first 12 frames only (slow)
        GL.set_map("order_limit", order)
        if order < 3:
            seq = R.generate_sequential_teps(64, order)
            res[f"seq{order}"] = np.packbits(np.concatenate([np.asarray(s) for s in seq], 0).astype(np.uint8), axis=1)
        S, NT = [], []
        for i in range(len(y) if order < 3 else 12):
            log = f"./log/FS-OSD-order-{order}.txt"
            if os.path.exists(log):
                os.remove(log)
            with _quiet():
                R.fs_osd(2.5, 0.1, DS(y[i], lab[i]))
            txt = open(log).read()
            m = re.search(r"S:(\d+) F:(\d+)", txt)
            t = re.search(r"Average TEPs:([0-9.]+)", txt)
            S.append(int(m.group(1)))
            NT.append(int(round(float(t.group(1)))))
        res[f"success{order}"] = np.array(S)
        res[f"num_teps{order}"] = np.array(NT)
    np.savez_compressed(os.path.join(GOLD, "fs_ref_shim.npz"), y=y, labels=lab.astype(np.uint8), beta=0.1, tau_psc=30, d_min=14, **res)


def run_pb():
    """PB_OSD: per-frame outcome of pb_osd (S/F, TEPs visited, improvement counters) parsed from its own log."""
    _enter("PB_OSD")
    import fill_matrix_info as F
    import globalmap as GL

    with _quiet():
        code = F.Code(os.path.join(REF, "PB_OSD", ALIST))
    GL.set_map("code_parameters", code)
    GL.set_map("termination_num_threshlod", 100)
    GL.set_map("convention_osd", False)
    GL.set_map("miracle_view", False)
    GL.set_map("pb_osd", True)
    import pb_testing as R

    class DS:
        def __init__(self, y, lab):
            self.b = [(y[None, :], lab[None, :])]

        def as_numpy_iterator(self):
            return iter(self.b)

    y, lab = _failed(code, 36, 5)
    res = {}
    for order, snr in ((1, 2.5), (2, 2.5), (2, 3.5), (3, 2.5)):  # order_limit 3 is the reference's default (PB_OSD/globalmap.py:42)
        GL.set_map("order_limit", order)
        S, NT, ML, A1, A2 = [], [], [], [], []
        for i in range(len(y)):
            log = f"./log/PB-OSD-order-{order}.txt"
            if os.path.exists(log):
                os.remove(log)
            with _quiet():
                R.pb_osd(snr, DS(y[i], lab[i]))
            txt = open(log).read()
            m = re.search(r"S/F:(\d+)/(\d+)", txt)
            t = re.search(r"Average TEPs:([0-9.]+) Maintained_list_len:([0-9.]+) Average_suc: ([0-9.]+)/([0-9.]+)", txt)
            S.append(int(m.group(1)))
            NT.append(int(round(float(t.group(1)))))
            ML.append(int(round(float(t.group(2)))))
            A1.append(int(round(float(t.group(3)))))
            A2.append(int(round(float(t.group(4)))))
        tag = f"o{order}_snr{int(snr * 10)}"
        res[f"success_{tag}"] = np.array(S)
        res[f"num_teps_{tag}"] = np.array(NT)
        res[f"list_cmp_{tag}"] = np.array(ML)
        res[f"suc1_{tag}"] = np.array(A1)
        res[f"suc2_{tag}"] = np.array(A2)
        print(tag, S, NT, flush=True)
    np.savez_compressed(os.path.join(GOLD, "pb_ref_shim.npz"), y=y, labels=lab.astype(np.uint8), **res)


def run_dl():
    """DL_OSD_Testing_serial: H-based ordering, TEP blocks, sliding_osd with a fixed window classifier."""
    _enter("DL_OSD_Testing_serial")
    import fill_matrix_info as F
    import globalmap as GL
    import tensorflow as tf

    with _quiet():
        code = F.Code(os.path.join(REF, "DL_OSD_Testing_serial", ALIST))
    for k, v in dict(code_parameters=code, num_iterations=12, threshold_sum=2, segment_num=6, soft_margin=0.9, decoding_length=30,
                     sliding_win_width=5, convention_path=True).items():
        GL.set_map(k, v)
    import ordered_statistics_decoding as R

    sys.path.insert(0, ROOT)
    from oracle import nms_oracle as NO

    y, lab = _failed(code, 12, 4)
    traj = NO.decode(y, code.H, 12)["traj"]  # [B,13,128], row 0 = channel
    rng = np.random.default_rng(1234)
    taps = (np.full(13, 1 / 13) + 0.05 * rng.normal(size=13)).astype(np.float32)
    new_inputs = (np.einsum("bij,i->bj", traj.astype(np.float64), taps.astype(np.float64))).astype(np.float32)
    osd = R.osd(code)
    _, boundary = GL.secure_segment_threshold()
    ranges = [range(boundary[i], boundary[i + 1]) for i in range(6)]
    path = [[0] * 6]
    for a in range(6):
        p = [0] * 6
        p[a] = 1
        path.append(p)
    for a in range(6):
        for b in range(a, 6):
            p = [0] * 6
            p[a] += 1
            p[b] += 1
            if p[a] <= len(ranges[a]) and p[b] <= len(ranges[b]):
                path.append(p)
    blocks = [osd.error_pattern_gen(p, ranges) for p in path]
    acc = np.insert(np.cumsum([b.shape[0] for b in blocks]), 0, 0)
    # window classifier with a readable rule: stop when (2nd smallest - smallest) of the window is large,
    # more readily at deeper positions k:  logit1 - logit0 = 2*(w[1]-w[0]) - 2.5 + 0.2*k
    W = np.eye(6, dtype=np.float32)
    V = np.zeros((6, 2), dtype=np.float32)
    V[0, 1], V[1, 1], V[5, 1] = -2.0, 2.0, 0.2
    V[:, 0] = 0.0
    bias_trick = np.float32(2.5)

    def fcn(x):
        h = np.asarray(x, dtype=np.float32) @ W
        o = h @ V
        o[..., 0] += bias_trick
        e = np.exp(o - o.max(axis=-1, keepdims=True))
        return tf.constant(e / e.sum(axis=-1, keepdims=True))

    input_list = traj.reshape(-1, 128)
    spy = {"mins": []}
    real = osd.acquire_min

    def acquire(*a, **k):
        v = real(*a, **k)
        spy["mins"].append(float(np.asarray(v)))
        return v

    osd.acquire_min = acquire
    per, mins = [], []
    idx_lists, Ms = [], []
    order_H, order_in, order_orig, order_lab = osd.check_matrix_reorder(input_list, new_inputs, lab)
    with _quiet():
        upd_idx, upd_M, swap_len, swap_pos = osd.identify_mrb(np.asarray(order_H))
    lri = np.asarray(osd.mag_input_gen(new_inputs))
    for i in range(len(y)):
        spy["mins"] = []
        with _quiet():
            s, f, w, c = osd.sliding_osd(fcn, input_list[13 * i:13 * i + 13], new_inputs[i:i + 1], lab[i:i + 1], (blocks, acc))
        per.append((s, f, w, int(c)))
        mins.append(spy["mins"] + [np.nan] * (len(blocks) - len(spy["mins"])))
        idx_lists.append(np.asarray(upd_idx[i]))
        Ms.append(np.packbits(np.asarray(upd_M[i]).astype(np.uint8), axis=1))
    np.savez_compressed(
        os.path.join(GOLD, "dl_ref_shim.npz"), y=y, labels=lab.astype(np.uint8), traj=traj, taps=taps, new_inputs=new_inputs,
        path=np.array(path), block_sizes=np.array([b.shape[0] for b in blocks]),
        blocks=np.packbits(np.concatenate(blocks, 0).astype(np.uint8), axis=1), boundary=np.asarray(boundary), W=W, V=V,
        per_frame=np.array(per), block_mins_fp32=np.array(mins), fcn_bias0=bias_trick, lri=lri, upd_idx=np.array(idx_lists), M=np.array(Ms), swap_len=np.array(swap_len))


def run_gf2():
    """full_gf2elim lifted with ast from PB_OSD/pb_testing.py (pure NumPy inside a TF-importing module)."""
    src = open(os.path.join(REF, "PB_OSD", "pb_testing.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "full_gf2elim")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "pb_testing.py", "exec"), ns)
    G = np.load(os.path.join(GOLD, "code_ref.npz"))["G"].astype(np.int64)
    H = np.load(os.path.join(GOLD, "code_ref.npz"))["H"].astype(np.int64)
    rng = np.random.default_rng(7)
    perms, mats, swaps, nsw = [], [], [], []
    for t in range(60):
        p = rng.permutation(128)
        A = (G if t % 2 == 0 else H)[:, p].copy()
        R, sw = ns["full_gf2elim"](A)
        perms.append(p)
        mats.append(np.packbits(R.astype(np.uint8), axis=1))
        s = np.full((40, 2), -1)
        s[:len(sw)] = np.array(sw).reshape(-1, 2)
        swaps.append(s)
        nsw.append(len(sw))
    np.savez_compressed(os.path.join(GOLD, "gf2elim_ref.npz"), perms=np.array(perms), reduced=np.array(mats), swaps=np.array(swaps),
                        n_swaps=np.array(nsw))


CMDS = {"gen": run_gen, "nms": run_nms, "osd": run_osd, "fs": run_fs, "pb": run_pb, "dl": run_dl, "gf2": run_gf2}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if "--real-tf" in sys.argv:
        REAL_TF = True
    if "--out" in sys.argv:
        GOLD = os.path.abspath(sys.argv[sys.argv.index("--out") + 1])
    os.makedirs(GOLD, exist_ok=True)
    if which == "all":
        for name in ("gen", "gf2", "nms", "osd", "fs", "pb", "dl"):
            print("==", name, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), name] + sys.argv[2:], check=True)
    else:
        CMDS[which]()
