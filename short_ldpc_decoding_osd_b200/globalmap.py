"""Process-global configuration dictionary with the reference's accessors.

Mirror of the reference's ``globalmap.py`` (eight copies, e.g. ``LDPC_128/Ldpc_128_testing/globalmap.py:8-22``,
``LDPC_128/PB_OSD/globalmap.py:26-47``, ``LDPC_128/FS_OSD/globalmap.py``, ``LDPC_128/DL_OSD_Testing_serial/globalmap.py:27-76``):
``set_map / get_map / del_map`` keep their behaviour (a missing key prints a message and returns None), and
``global_setting(argv)`` sets the union of the keys the four testing drivers set, with the reference's values.
"""
from __future__ import annotations

import numpy as np

from . import fill_matrix_info as Fill_matrix

map = {}  # noqa: A001  (reference name)


def set_map(key, value):
    map[key] = value


def del_map(key):
    try:
        del map[key]
    except KeyError:
        print("key:'" + str(key) + "' non-existence")


def get_map(key):
    try:
        if key in "all":
            return map
        return map[key]
    except KeyError:
        print("key:'" + str(key) + "' non-existence")


def global_setting(argv):
    """argv = "python snr_lo snr_hi snr_num unit_batch_size num_iterations H_filename decoder_type".split()
    (PB_OSD/Main_PB_OSD.py:13, FS_OSD/Main_FS_OSD.py:13, DL_OSD_Testing_serial/Main_DL_OSD.py)."""
    set_map("snr_lo", float(argv[1]))
    set_map("snr_hi", float(argv[2]))
    set_map("snr_num", int(argv[3]))
    set_map("unit_batch_size", int(argv[4]))
    set_map("num_iterations", int(argv[5]))
    set_map("H_filename", argv[6])
    set_map("selected_decoder_type", argv[7])
    set_map("ALL_ZEROS_CODEWORD_TRAINING", False)
    import os

    path = argv[6] if os.path.exists(argv[6]) else Fill_matrix.CCSDS_ALIST
    set_map("code_parameters", Fill_matrix.Code(path))
    # PB_OSD/globalmap.py:42-47, FS_OSD/globalmap.py:44-50
    set_map("order_limit", 3)
    set_map("termination_num_threshlod", 100)
    set_map("convention_osd", False)
    set_map("miracle_view", False)
    set_map("pb_osd", True)
    set_map("fs_osd", True)
    set_map("d_min", 14)
    set_map("tau_psc", 30)
    # DL_OSD_Testing_serial/globalmap.py:45-55
    set_map("print_interval", 100)
    set_map("record_interval", 100)
    set_map("convention_path", False)
    set_map("termination_threshold", 500)
    set_map("threshold_sum", 3)
    set_map("training_snr", 2.7)
    set_map("segment_num", 6)
    set_map("soft_margin", 0.9)
    set_map("decoding_length", 30)
    set_map("sliding_win_width", 5)


def secure_segment_threshold():
    """DL_OSD_Testing_serial/globalmap.py:57-76 -> (segment sizes [1,4,8,12,16,23], boundaries [0,1,5,13,25,41,64])."""
    num_seg = get_map("segment_num")
    code = get_map("code_parameters")
    allocation_length = code.k - 1
    basic_length = list(range(1, num_seg))
    num_basic = sum(basic_length)
    sizes = [int(allocation_length / num_basic * b) for b in basic_length]
    sizes[-1] += allocation_length - sum(sizes)
    whole = np.insert(sizes, 0, 1)
    return whole, np.insert(np.cumsum(whole), 0, 0)
