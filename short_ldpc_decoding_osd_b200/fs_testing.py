"""Drop-in for ``LDPC_128/FS_OSD/fs_testing.py``: fast-and-scalable OSD (Choi & Jeong 2019 as re-implemented
by the reference).

* ``generate_sequential_teps(max_value, max_loops)`` -> list of ``int32[C(64,w),64]`` per weight (``:32-49``)
* ``acquire_pnc_boundary(descending_input)`` (``:22-30``)
* ``one_tep_compare(updated_inputs, nth_tep, reduced_G, threshold)`` (``:51-64``)
* ``swapped_info`` / ``full_gf2elim`` / ``identify_mrb`` (``:233-322``, same code as PB_OSD)
* ``fs_osd(snr, beta, selected_ds)`` (``:68-231``): the per-frame policy (order-0 acceptance below tau_e,
  order-skip rule with beta, per-TEP tau_e stop that does not update the decision, tau_psc-gated improvement)
  runs for ALL frames of the dataset in one call of ldpcb_osd_fs_decode_host; the reference's log file
  ``./log/FS-OSD-order-p.txt`` is written with the same lines.
"""
from __future__ import annotations

import math
import os
import time
from collections import Counter

import numpy as np

from . import _lib
from . import convention_osd as cnv_OSD
from . import globalmap as GL
from .pb_testing import _frames_of, full_gf2elim, identify_mrb, miracle_view, swapped_info, swapped_info_batch  # noqa: F401
from .runtime import get_handle


def acquire_pnc_boundary(descending_input):
    order_limit = GL.get_map("order_limit")
    k = GL.get_map("code_parameters").k
    a = np.abs(np.asarray(descending_input, dtype=np.float32))
    return [np.float32(sum(a[k - (i + 1):k])) for i in range(order_limit)]


def generate_sequential_teps(max_value, max_loops):
    h = get_handle()
    words = h.tep_table(int(max_loops), _lib.TEP_FS)
    bounds = [1, 65, 2081, 43745]
    mats = cnv_OSD.unpack_tep_words(words, max_value)
    return [mats[bounds[w - 1]:bounds[w]] for w in range(1, max_loops + 1)]


def one_tep_compare(updated_inputs, nth_tep, reduced_G, threshold):
    """(early_stopping, optimal_codeword int32[1,128], w_dmin) for one TEP on permuted inputs (:51-64).
    The weighted distance is returned as a float computed from the exact integer score."""
    yo = np.ascontiguousarray(np.asarray(updated_inputs, dtype=np.float32).reshape(1, 128))
    words = np.ascontiguousarray(cnv_OSD._redG_words(reduced_G).reshape(1, 64))
    teps = cnv_OSD.pack_tep_matrix(np.asarray(nth_tep).reshape(1, 64))
    h = get_handle()
    cw = np.empty((1, 4), np.uint32)
    bq = np.empty(1, np.int64)
    ex = np.empty(1, np.int32)
    h.call("ldpcb_osd_sweep_host", yo, yo, words, 1, teps, 1, 0, cw, None, bq, ex)
    cand = _lib.unpack_bits(cw).astype(np.int32)
    hard = np.where(yo > 0, 0, 1)
    early = float((cand ^ hard).sum()) < threshold
    return early, cand, float(np.ldexp(float(bq[0]), int(ex[0]) - 54))


def fs_osd_batch(inputs, labels, order_limit, beta, tau_e=None, tau_psc=None):
    """The FS policy on [B,128] channel LLRs -> dict(correct, num_teps, stop_kind, best_tep, codeword)."""
    y = np.ascontiguousarray(np.asarray(inputs, dtype=np.float32).reshape(-1, 128))
    B = y.shape[0]
    code = GL.get_map("code_parameters")
    if tau_e is None:
        tau_e = math.floor(GL.get_map("d_min") - 1) / 2  # fs_testing.py:92 (precedence as written: 6.5 for d_min 14)
    if tau_psc is None:
        tau_psc = GL.get_map("tau_psc")
    shift = float(np.float32(beta * (code.check_matrix_column - code.k)))
    h = get_handle()
    cw = np.empty((B, 4), np.uint32)
    best = np.empty(B, np.int32)
    num = np.empty(B, np.int32)
    kind = np.empty(B, np.uint8)
    h.call("ldpcb_osd_fs_decode_host", y, B, int(order_limit), float(tau_e), int(tau_psc), shift, cw, best, num, kind)
    codeword = _lib.unpack_bits(cw)
    correct = (codeword == (np.asarray(labels).reshape(B, 128) & 1)).all(axis=1)
    return {"correct": correct, "num_teps": num, "stop_kind": kind, "best_tep": best, "codeword": codeword}


def fs_osd(snr, beta, selected_ds):
    start_time = time.process_time()
    order_limit = GL.get_map("order_limit")
    y, lab = _frames_of(selected_ds)
    logdir = "./log/"
    os.makedirs(logdir, exist_ok=True)
    limit = GL.get_map("termination_num_threshlod") or 100
    summary = {"snr": snr, "order_limit": order_limit, "frames": len(y)}
    if GL.get_map("miracle_view"):
        from .pb_testing import pb_osd

        return pb_osd(snr, list(zip(y[:, None, :], lab[:, None, :])))
    if GL.get_map("convention_osd"):
        res = cnv_OSD.convention_osd_batch(y, lab, order_limit)
        ok = res["correct"]
        S, F = int(ok.sum()), int((~ok).sum())
        counter = Counter(int(p) for p in res["phase"])
        FER = round(F / max(S + F, 1), 4)
        log_filename = logdir + "CNV-OSD-order-" + str(order_limit) + ".txt"
        T2 = time.process_time()
        print("\nFor Conv-OSD %.1fdB (order_limit:%d) :\n" % (snr, order_limit))
        print("----> S:" + str(S) + " F:" + str(F) + "\n")
        print("Distribution of phases:" + str(counter) + "\n")
        print("FER:" + str(FER) + " Average TEPs size:", res["teps_size"], "\n")
        with open(log_filename, "a+") as f:
            f.write("\nFor CNV-OSD %.1fdB (order_limit:%d) summary:\n" % (snr, order_limit))
            f.write("----> S:" + str(S) + " F:" + str(F) + "\n")
            f.write("Distribution of phases:" + str(counter) + "\n")
            f.write("FER:" + str(FER) + " Average TEPs size:" + str(res["teps_size"]) + "\n")
            f.write(f"Running time:{T2 - start_time} seconds with mean time {(T2 - start_time)/max(S + F, 1):.4f}!")
        summary.update({"S": S, "F": F, "FER": FER, "phases": dict(counter), "log": log_filename})
        return summary
    res = fs_osd_batch(y, lab, order_limit, beta)
    fails_cum = np.cumsum(~res["correct"])
    n_used = int(np.searchsorted(fails_cum, limit) + 1) if fails_cum.size and fails_cum[-1] >= limit else len(y)
    ok = res["correct"][:n_used]
    correct_sum, fail_sum = int(ok.sum()), int((~ok).sum())
    total_num = correct_sum + fail_sum
    counter_teps_sum = int(res["num_teps"][:n_used].sum())
    FER = round(fail_sum / max(total_num, 1), 4)
    average_size = round(counter_teps_sum / max(total_num, 1), 5)
    log_filename = logdir + "FS-OSD-order-" + str(order_limit) + ".txt"
    T2 = time.process_time()
    print("\nFor FS-OSD %.1fdB (order_limit:%d) :\n" % (snr, order_limit))
    print("----> S:" + str(correct_sum) + " F:" + str(fail_sum) + "\n")
    print(f"FER:{FER:.2f} Average TEPs:{average_size:.2f} \n")
    with open(log_filename, "a+") as f:
        f.write("\nFor FS-OSD %.1fdB (order_limit:%d) summary:\n" % (snr, order_limit))
        f.write("----> S:" + str(correct_sum) + " F:" + str(fail_sum) + "\n")
        f.write(f"FER:{FER:.2f} Average TEPs:{average_size:.2f}\n")
        f.write(f"Running time:{T2 - start_time} seconds with mean time {(T2 - start_time)/max(total_num, 1):.4f}!\n")
    summary.update({"S": correct_sum, "F": fail_sum, "FER": FER, "average_teps": average_size, "log": log_filename,
                    "stop_kinds": dict(Counter(int(k) for k in res["stop_kind"][:n_used]))})
    return summary
