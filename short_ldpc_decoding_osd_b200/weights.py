"""Trained weights and the learned decoding path as plain files.

The reference keeps them in TF checkpoints and a pickle (ldpc_128_testing.py:57-68, nn_testing.py:38-64,84-107);
``scripts/export_tf_weights.py`` converts those where TensorFlow exists.  This module is the TF-free consumer:

    w = weights.load_npz("weights.npz")
    model.layer.shared_check_weight[:] = w["nms_check"]            # ms_test.Decoding_model
    nn = weights.make_conv_bitwise(w); fcn = weights.make_fcn(w)   # nn_net.conv_bitwise / Predict_outlier_light
    path = weights.load_decoding_path("decoding_path.json")        # -> nn_testing.filter_order_patterns(path)
"""
from __future__ import annotations

import json
from typing import Dict, List

import numpy as np

SHAPES = {"nms_check": (1,), "k1": (3, 1, 8), "k2": (3, 8, 4), "k3": (3, 4, 2), "dense_w": (14, 1), "dense_b": (1,)}


def load_npz(path: str) -> Dict[str, np.ndarray]:
    """Arrays of an exported weights file, float32, shapes checked against the reference's layers."""
    with np.load(path) as z:
        w = {k: np.asarray(z[k], dtype=np.float32) for k in z.files}
    for k, shp in SHAPES.items():
        if k in w and tuple(w[k].shape) != shp:
            raise ValueError(f"{path}: {k} has shape {tuple(w[k].shape)}, the reference's layer has {shp}")
    if "fcn1" in w:
        n = w["fcn1"].shape[0]
        if w["fcn1"].shape != (n, n) or w.get("fcn2", np.zeros((n, 2))).shape != (n, 2):
            raise ValueError(f"{path}: fcn1/fcn2 must be [w+1,w+1] and [w+1,2] (nn_net.py:140-144)")
    return w


def alpha_of(w: Dict[str, np.ndarray]) -> float:
    """softplus of the raw NMS check weight (ms_test.py:207-208): the alpha_check argument of ldpcb_nms_decode."""
    from .runtime import softplus

    return softplus(np.asarray(w["nms_check"]).reshape(-1)[0])


def make_conv_bitwise(w: Dict[str, np.ndarray]):
    from . import nn_net

    nn = nn_net.conv_bitwise()
    nn.set_weights(w["k1"], w["k2"], w["k3"], w["dense_w"], w["dense_b"])
    return nn


def make_fcn(w: Dict[str, np.ndarray]):
    from . import nn_net

    return nn_net.Predict_outlier_light(w["fcn1"].shape[0] - 1, w["fcn1"], w["fcn2"])


def load_decoding_path(path: str) -> List[List[int]]:
    """Order patterns in the learned visiting order (descending training count), as query_decoding_path returns
    them before filter_order_patterns (nn_testing.py:84-114)."""
    with open(path) as fh:
        d = json.load(fh)
    out = [[int(x) for x in p] for p in d["decoding_path"]]
    if any(len(p) != len(out[0]) for p in out):
        raise ValueError(f"{path}: order patterns of unequal length")
    return out
