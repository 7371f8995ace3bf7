#!/usr/bin/env python
"""SURVEY.md 8(f) row f4 -- run this WHERE TENSORFLOW EXISTS, inside the reference tree, to turn the reference's TF
checkpoints and decoding-path pickle into the plain files this repo consumes (weights.npz + decoding_path.json).

    cd <reference>/LDPC_128/DL_OSD_Testing_serial        # or Ldpc_128_testing for the NMS weight alone
    python <this repo>/scripts/export_tf_weights.py --out weights.npz \
        [--nms-ckpt ../Ldpc_128_training/ckpts/NMS-1/12th/] \
        [--cnn-ckpt <dir of the conv_bitwise checkpoints>] [--fcn-ckpt <dir of the Predict_outlier_light checkpoints>] \
        [--path-pickle ../DL_Training_serial/log/NMS-1/2.7-2.7dB/dist-error-pattern-benchmark.pkl --path-out decoding_path.json]

What it restates: tf.train.Checkpoint(myAwesomeModel=model).restore(latest) exactly as
Ldpc_128_testing/ldpc_128_testing.py:57-68 and DL_OSD_Testing_serial/nn_testing.py:38-64 do, through the
reference's own model classes (imported from the current directory), then reads the variables:
    nms_check  raw check weight, before softplus                 ms_test.py:83
    k1 k2 k3   Conv1D kernels [3,1,8], [3,8,4], [3,4,2]          nn_net.py:185-188
    dense_w dense_b   Dense(1) kernel [14,1] and bias [1]        nn_net.py:189
    fcn1 fcn2  Dense(6, no bias), Dense(2, no bias)              nn_net.py:140-144
The decoding path is the sixth object of the pickle, a dict "order pattern" -> count, sorted by descending count
(nn_testing.py:84-107).  This script cannot be exercised in the build container (no TensorFlow); its consumer,
short_ldpc_decoding_osd_b200.weights, is tested with synthetic files of the same layout.
"""
import argparse, json, os, pickle, re, sys

import numpy as np


def restore(model, ckpt_dir, tf):
    ckpt = tf.train.latest_checkpoint(ckpt_dir)
    if ckpt is None:
        raise SystemExit(f"no checkpoint under {ckpt_dir}")
    tf.train.Checkpoint(myAwesomeModel=model).restore(ckpt).expect_partial()
    return ckpt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="weights.npz")
    ap.add_argument("--nms-ckpt")
    ap.add_argument("--cnn-ckpt")
    ap.add_argument("--fcn-ckpt")
    ap.add_argument("--path-pickle")
    ap.add_argument("--path-out", default="decoding_path.json")
    ap.add_argument("--argv", default="python 2.0 3.0 6 100  12 CCSDS_ldpc_n128_k64.alist NMS-1",
                    help="the argv string the reference's driver passes to GL.global_setting (Main_DL_OSD.py:13)")
    a = ap.parse_args()
    sys.path.insert(0, os.getcwd())
    out = {}
    if a.nms_ckpt or a.cnn_ckpt or a.fcn_ckpt:
        import tensorflow as tf
        import globalmap as GL
        import fill_matrix_info as Fill_matrix

        # the drivers set their globals from a fixed argv (Main_DL_OSD.py:13-16, ldpc_128_testing.py likewise)
        GL.global_setting(a.argv.split())
        code = Fill_matrix.Code(GL.get_map("H_filename"))
        GL.set_map("code_parameters", code)
        if a.nms_ckpt:
            import ms_test as Decoder_module

            model = Decoder_module.Decoding_model()
            y = tf.zeros([1, code.check_matrix_column], tf.float32)
            model(y, tf.zeros([1, code.check_matrix_column], tf.int64))  # builds the variables
            print("NMS:", restore(model, a.nms_ckpt, tf))
            out["nms_check"] = np.asarray(model.layer.shared_check_weight.numpy(), np.float32)
        if a.cnn_ckpt or a.fcn_ckpt:
            import nn_net as CRNN_DEF

            if a.cnn_ckpt:
                nn = CRNN_DEF.conv_bitwise()
                nn(tf.zeros([code.check_matrix_column, GL.get_map("num_iterations") + 1, 1], tf.float32))
                print("CNN:", restore(nn, a.cnn_ckpt, tf))
                out.update(k1=nn.cnv_one.kernel.numpy(), k2=nn.cnv_two.kernel.numpy(), k3=nn.cnv_three.kernel.numpy(),
                           dense_w=nn.dense.kernel.numpy(), dense_b=nn.dense.bias.numpy())
            if a.fcn_ckpt:
                w = GL.get_map("sliding_win_width")
                fcn = CRNN_DEF.Predict_outlier_light(w)
                fcn(tf.zeros([1, w + 1], tf.float32))
                print("FCN:", restore(fcn, a.fcn_ckpt, tf))
                out.update(fcn1=fcn.dense1.kernel.numpy(), fcn2=fcn.dense2.kernel.numpy())
        np.savez(a.out, **out)
        print("wrote", a.out, sorted(out))
    if a.path_pickle:
        with open(a.path_pickle, "rb") as fh:
            for _ in range(5):
                pickle.load(fh)
            pattern_dict = pickle.load(fh)
        ordered = sorted(pattern_dict, key=pattern_dict.get, reverse=True)
        path = [[int(t) for t in re.findall(r"\w+", s)] for s in ordered]
        with open(a.path_out, "w") as fh:
            json.dump({"decoding_path": path, "counts": [int(pattern_dict[s]) for s in ordered]}, fh)
        print("wrote", a.path_out, len(path), "order patterns")


if __name__ == "__main__":
    main()
