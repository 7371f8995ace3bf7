"""Drop-in for the driver side of ``LDPC_128/DL_OSD_Testing_serial/nn_testing.py``.

* ``query_convention_path`` (``:66-82``), ``filter_order_patterns`` (``:116-123``), ``generate_teps`` (``:144-157``)
* ``Testing_OSD(snr, selected_ds, restore_list, indicator_list, prefix_list, DIA, nn=None, fcn=None)``
  (``:159-256``): per batch the DIA FIR, the block minima of every frame and the window policy; same log file
  and return value ``(FER, log_filename)``.  Checkpoint restoring (``NN_gen``, ``:38-64``) is replaced by
  passing the two networks (see nn_net.py) because TF checkpoints cannot be read without TensorFlow; the
  learned decoding path pickle (``query_decoding_path``, ``:84-114``) is read if present.
"""
from __future__ import annotations

import os
import pickle
import re
import time

import numpy as np

from . import globalmap as GL
from . import nn_net
from . import ordered_statistics_decoding as OSD_mod


def query_convention_path(indicator_list=None, prefix_list=None, DIA=False):
    nn_type = "benchmark"
    order_sum = GL.get_map("threshold_sum") + 1
    if DIA and indicator_list:
        for i, element in enumerate(indicator_list):
            if element:
                nn_type = prefix_list[i]
                break
    decoding_path, seen = [], set()
    for i in range(order_sum):
        for j1 in range(order_sum):
            for j2 in range(order_sum):
                for j3 in range(order_sum):
                    if j1 + j2 + j3 <= i and (j1, j2, j3) not in seen:  # tf.raw_ops.UniqueV2 keeps first occurrences
                        seen.add((j1, j2, j3))
                        decoding_path.append([j1, j2, j3])
    return decoding_path, nn_type


def filter_order_patterns(decoding_path):
    threshold_sum = GL.get_map("threshold_sum")
    length = GL.get_map("decoding_length")
    return [p for p in decoding_path if sum(p) <= threshold_sum][:length]


def query_decoding_path(indicator_list, prefix_list, DIA, log_dir=None):
    nn_type = "benchmark"
    if DIA:
        for i, element in enumerate(indicator_list):
            if element:
                nn_type = prefix_list[i]
                break
    decoder_type = GL.get_map("selected_decoder_type")
    log_dir = log_dir or "../DL_Training_serial/log/" + decoder_type + "/2.7-2.7dB/"
    with open(log_dir + "dist-error-pattern-" + nn_type + ".pkl", "rb") as fh:
        for _ in range(5):
            pickle.load(fh)
        pattern_dict = pickle.load(fh)
    string_pattern = sorted(pattern_dict, key=pattern_dict.get, reverse=True)
    decoding_path = [[int(d) for d in re.findall(r"\w+", s)] for s in string_pattern]
    return filter_order_patterns(decoding_path), nn_type


def convention_segment_path(threshold_sum=None, num_segments=None):
    """All order patterns over the DL segments with total weight <= threshold_sum, by increasing weight
    (the exhaustive path used when no learned path is available; 84 patterns for 6 segments, sum <= 3)."""
    threshold_sum = GL.get_map("threshold_sum") if threshold_sum is None else threshold_sum
    num_segments = GL.get_map("segment_num") if num_segments is None else num_segments
    sizes, _ = GL.secure_segment_threshold()
    out = []

    def rec(prefix, left):
        if len(prefix) == num_segments:
            out.append(list(prefix))
            return
        for v in range(0, min(left, int(sizes[len(prefix)])) + 1):
            rec(prefix + [v], left - v)

    rec([], threshold_sum)
    out.sort(key=lambda p: (sum(p), [-x for x in p]))
    return out


def generate_teps(osd, residual_path):
    num_segments = GL.get_map("segment_num")
    _, boundary_MRB = GL.secure_segment_threshold()
    range_list = [range(boundary_MRB[i], boundary_MRB[i + 1]) for i in range(num_segments)]
    error_pattern_list, sizes = [], []
    for path in residual_path:
        element = osd.error_pattern_gen(path, range_list)
        error_pattern_list.append(element)
        sizes.append(element.shape[0])
    return error_pattern_list, np.insert(np.cumsum(sizes), 0, 0)


def Testing_OSD(snr, selected_ds, restore_list=None, indicator_list=None, prefix_list=None, DIA=True, nn=None, fcn=None,
                residual_path=None):
    start_time = time.process_time()
    code = GL.get_map("code_parameters")
    order_sum = GL.get_map("threshold_sum")
    soft_margin = GL.get_map("soft_margin")
    osd_instance = OSD_mod.osd(code)
    nn_type = "benchmark"
    if residual_path is None:
        if GL.get_map("convention_path"):
            residual_path = convention_segment_path()
        else:
            try:
                residual_path, nn_type = query_decoding_path(indicator_list or [True], prefix_list or ["model_cnn"], DIA)
            except (OSError, pickle.UnpicklingError):
                residual_path = filter_order_patterns(convention_segment_path())
    tep_info = generate_teps(osd_instance, residual_path)
    list_length = GL.get_map("num_iterations") + 1
    if DIA and nn is None:
        nn = nn_net.conv_bitwise()
    if fcn is None:
        fcn = nn_net.Predict_outlier_light(GL.get_map("sliding_win_width"))
    logdir = "./log/"
    os.makedirs(logdir, exist_ok=True)
    log_filename = logdir + "OSD-" + str(order_sum) + "-" + nn_type + ".txt"
    it = selected_ds.as_numpy_iterator() if hasattr(selected_ds, "as_numpy_iterator") else iter(selected_ds)
    fail_sum = correct_sum = windows_sum = complexity_sum = actual_size = 0
    for batch in it:
        inputs_all = np.asarray(batch[0], dtype=np.float32)
        labels_all = np.asarray(batch[1])
        if DIA:
            squashed, _, labels = nn.preprocessing_inputs((inputs_all, labels_all))
            new_inputs = nn(squashed)
        else:
            labels = labels_all[0::list_length]
            new_inputs = inputs_all[0::list_length]
        actual_size += labels.shape[0]
        c, f, w, cx = osd_instance.sliding_osd(fcn, inputs_all, new_inputs, labels, tep_info)
        correct_sum += c
        fail_sum += f
        windows_sum += w
        complexity_sum += cx
        if fail_sum >= GL.get_map("termination_threshold"):
            break
    T2 = time.process_time()
    FER = round(fail_sum / max(actual_size, 1), 5)
    average_size = round(complexity_sum / max(actual_size, 1), 4)
    wins_size = round(windows_sum / max(actual_size, 1), 4)
    print("\nFor %.1fdB (order_sum:%d) " % (snr, order_sum) + nn_type + ":\n")
    print("----> S:" + str(correct_sum) + " F:" + str(fail_sum) + "\n")
    print(f"FER:{FER}--> S/F:{correct_sum} /{fail_sum} Avr TEPs:{average_size} Wins:{wins_size}")
    with open(log_filename, "a+") as f:
        f.write(f"For {snr:.1f}dB order_sum:{order_sum} len:{len(residual_path)} soft_margin:{soft_margin}:\n")
        f.write(f"Selected actual path:{residual_path}\n")
        f.write("----> S:" + str(correct_sum) + " F:" + str(fail_sum) + "\n")
        f.write(f"FER:{FER}--> S/F:{correct_sum} /{fail_sum} Avr TEPs:{average_size} Wins:{wins_size}\n")
        f.write(f"Running time:{T2 - start_time} seconds with mean time {(T2 - start_time)/max(actual_size, 1):.4f}!\n")
    return FER, log_filename
