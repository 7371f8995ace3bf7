"""Reference arm under a real TensorFlow (TEST / BASELINE INFRASTRUCTURE, never imported by the product).

Times the UNMODIFIED reference sources for bench.py's workload, as BASELINE.md section 3 step 1 asks:

  stage nms   LDPC_128/Ldpc_128_testing/ms_test.py  Decoding_model()(inputs, labels) in batches of
              unit_batch_size = 1000 (ldpc_128_testing.py:20,119)
  stage osd   LDPC_128/FS_OSD/fs_testing.py swapped_info (:308-322) followed by
              LDPC_128/FS_OSD/convention_osd.py convention_osd_main (:49-77) for every frame the NMS stage left with
              a non-zero syndrome, on the channel LLR (pb_testing.py:71-72), TEP matrix generated once

The two stages run in separate child processes because the reference's directories define modules of the same
names (globalmap, fill_matrix_info).  Frames are the first `--frames` frames of the GPU arm's run (Philox seed
2024).  Prints one JSON object.  `--root` is a copy of the reference repository (baseline/_ref or
$LDPCB_REFERENCE_ROOT): nothing here reads /root/reference.

    python oracle/tf_reference_arm.py --root baseline/_ref --frames 500 --order 2 --steps 3 --warmup 1
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ALIST = "CCSDS_ldpc_n128_k64.alist"


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def stage_nms(a):
    d = os.path.join(a.root, "LDPC_128", "Ldpc_128_testing")
    sys.path.insert(0, d)
    os.chdir(tempfile.mkdtemp(prefix="tfarm_"))
    import fill_matrix_info as F
    import globalmap as GL

    with _quiet():
        code = F.Code(os.path.join(d, ALIST))
    GL.set_map("selected_decoder_type", "NMS-1")
    GL.set_map("num_iterations", 12)
    GL.set_map("code_parameters", code)
    import ms_test as R  # the reference module, unmodified
    import tensorflow as tf

    z = np.load(a.data)
    y, lab = z["y"], z["labels"].astype(np.int64)
    model = R.Decoding_model()
    times, index = [], None
    for s in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        idx = []
        for b0 in range(0, len(y), 1000):
            yb, lb = tf.constant(y[b0:b0 + 1000]), tf.constant(lab[b0:b0 + 1000])
            with _quiet():
                soft = model.layer(yb, lb)
                _, _, _, ind = model.get_eval(soft, lb)
                model.collect_failed_output_selective(soft, lb, ind)
            idx.append(np.asarray(ind).reshape(-1) + b0)
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            times.append(dt)
        index = np.concatenate(idx)
    np.savez(a.out, times=np.array(times), index=index, tf_version=str(getattr(tf, "__version__", "?")), tf_file=str(getattr(tf, "__file__", "?")))


def stage_osd(a):
    d = os.path.join(a.root, "LDPC_128", "FS_OSD")
    sys.path.insert(0, d)
    os.chdir(tempfile.mkdtemp(prefix="tfarm_"))
    import fill_matrix_info as F
    import globalmap as GL

    with _quiet():
        code = F.Code(os.path.join(d, ALIST))
    GL.set_map("code_parameters", code)
    GL.set_map("order_limit", a.order)
    import convention_osd as C
    import fs_testing as R
    import tensorflow as tf

    z = np.load(a.data)
    y, lab = z["y"], z["labels"].astype(np.int64)
    index = np.load(a.index)["index"]
    teps, bnd = C.generate_teps(a.order), C.query_boundary(a.order)
    times, ok = [], 0
    for s in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        ok = 0
        for i in index:
            ui, ul, rg = R.swapped_info(tf.constant(y[i]), tf.constant(lab[i]))
            good, _, _ = C.convention_osd_main((ui, ul, rg, teps, bnd))
            ok += bool(good)
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            times.append(dt)
    np.savez(a.out, times=np.array(times), ok=ok)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--root", required=True)
    p.add_argument("--frames", type=int, default=500)
    p.add_argument("--order", type=int, default=2)
    p.add_argument("--ebn0", type=float, default=2.5)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=1)
    p.add_argument("--stage", default="all")
    p.add_argument("--data")
    p.add_argument("--index")
    p.add_argument("--out")
    a = p.parse_args()
    if a.stage == "nms":
        return stage_nms(a)
    if a.stage == "osd":
        return stage_osd(a)
    sys.path.insert(0, ROOT)
    from oracle import philox_oracle as PO
    from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

    code = Code()
    y, cw, _ = PO.gen_frames(2024, 0, a.frames, a.ebn0, code.G)
    tmp = tempfile.mkdtemp(prefix="tfarm_")
    data, o1, o2 = os.path.join(tmp, "frames.npz"), os.path.join(tmp, "nms.npz"), os.path.join(tmp, "osd.npz")
    np.savez(data, y=y.astype(np.float32), labels=cw.astype(np.uint8))
    common = [sys.executable, os.path.abspath(__file__), "--root", a.root, "--order", str(a.order), "--steps", str(a.steps), "--warmup", str(a.warmup), "--data", data]
    subprocess.run(common + ["--stage", "nms", "--out", o1], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(common + ["--stage", "osd", "--index", o1, "--out", o2], check=True, stdout=subprocess.DEVNULL)
    n, o = np.load(o1), np.load(o2)
    t_nms, t_osd = n["times"], o["times"]
    nf = len(n["index"])
    secs = float(t_nms.sum() + t_osd.sum())
    shim = "tf_shim" in str(n["tf_file"])
    print(json.dumps({
        "value": a.frames * a.steps / secs, "ms_per_step": 1e3 * secs / a.steps, "frames_per_step": a.frames, "failed": nf,
        "kind": "tf-shim (NumPy emulation of TensorFlow, for testing this script only)" if shim else "tf",
        "tf_version": str(n["tf_version"]),
        "nms_frames_per_s": a.frames * a.steps / float(t_nms.sum()), "osd_frames_per_s": nf * a.steps / max(float(t_osd.sum()), 1e-9),
        "osd_ok": int(o["ok"]),
        "sample": f"{a.frames} frames/step x {a.steps} steps: the first {a.frames} frames of the GPU arm's run through the unmodified reference "
                  f"(Decoding_model B<=1000, then swapped_info + convention_osd_main order {a.order} on the {nf} detected failures)",
    }))


if __name__ == "__main__":
    main()
