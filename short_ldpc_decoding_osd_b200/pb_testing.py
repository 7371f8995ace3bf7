"""Drop-in for ``LDPC_128/PB_OSD/pb_testing.py`` (the parts on the hot path).

* ``swapped_info(inputs, labels)`` -> ``(updated_inputs f32[128], updated_labels int64[128], reduced_G int32[64,128])``
  -- pb_testing.py:306-320 (identical in fs_testing.py:308-322): reliability sort, GF(2) elimination and the
  pi2 permutation run in libldpc_b200.so (one frame = one warp of the OSD kernel at order 0).
* ``swapped_info_batch`` -- the same for [B,128] inputs in one call.
* ``full_gf2elim(M)`` -> ``(reduced M, recorded column swaps)`` -- pb_testing.py:231-266; host-side NumPy helper
  with the reference's pivot rule (the kernel does not call it; kept for API completeness and tests).
* ``identify_mrb(order_inputs, order_G)`` -- pb_testing.py:268-304, host-side, built on ``full_gf2elim``.
* ``miracle_view`` -- pb_testing.py:502-511, the genie statistic (MRB hard-decision errors per frame).
* ``pb_osd(snr, selected_ds)`` -- pb_testing.py:44-229 driver.  With ``GL.pb_osd`` set the PB policy (best-first
  TEP order, p_e^pro / p_e^suc stopping, :100-149,366-500) runs for all frames in one call of
  ldpcb_osd_pb_decode_host (order_limit 0..3; at order 3, the reference's default, the 43,745-entry TEP lists
  live in global memory); the ``convention_osd`` and ``miracle_view`` switches are served by the exhaustive GPU
  sweep.  Same log lines as the reference.
* ``pb_osd_batch(inputs, labels, snr, order_limit)`` -- the policy on [B,128] arrays.
"""
from __future__ import annotations

import os
import time
from collections import Counter

import numpy as np

from . import _lib
from . import convention_osd as cnv_OSD
from . import globalmap as GL
from .fill_matrix_info import gf2_systematic_form
from .runtime import get_handle


def full_gf2elim(M):
    return gf2_systematic_form(M)


def identify_mrb(order_inputs, order_G):
    code = GL.get_map("code_parameters")
    swapped_G, record = full_gf2elim(np.copy(order_G))
    index_order = np.arange(code.check_matrix_column)
    for a, b in record:
        index_order[a], index_order[b] = index_order[b], index_order[a]
    mrb = index_order[:code.k]
    mrb_swap = np.argsort(mrb, kind="stable")
    lrb = index_order[code.k:]
    lrb_swap = np.argsort(lrb, kind="stable")
    ident = np.identity(code.k, dtype=np.int32)
    interm = swapped_G[:, code.k:][:, lrb_swap]
    updated_lrb = (ident[:, mrb_swap].T.dot(interm) % 2).astype(np.int32)
    return np.concatenate([ident, updated_lrb], axis=1), np.concatenate([np.sort(mrb), np.sort(lrb)])


def _redG_matrix(words: np.ndarray) -> np.ndarray:
    """[...,64] uint64 P' words -> int32[...,64,128] reduced_G = [I | P']."""
    P = ((words[..., None] >> np.arange(64, dtype=np.uint64)) & np.uint64(1)).astype(np.int32)
    I = np.broadcast_to(np.identity(64, dtype=np.int32), P.shape)
    return np.concatenate([I, P], axis=-1)


def swapped_info_batch(inputs, labels):
    y = np.ascontiguousarray(np.asarray(inputs, dtype=np.float32).reshape(-1, 128))
    lab = np.asarray(labels).reshape(-1, 128)
    B = y.shape[0]
    h = get_handle()
    cw = np.empty((B, 4), np.uint32)
    perm = np.empty((B, 128), np.uint8)
    redg = np.empty((B, 64), np.uint64)
    h.call("ldpcb_osd_decode_host", y, y, B, 0, _lib.TEP_CONV, 0, cw, None, None, None, perm, redg)
    p = perm.astype(np.int64)
    return np.take_along_axis(y, p, axis=1), np.take_along_axis(lab, p, axis=1), _redG_matrix(redg), perm


def swapped_info(inputs, labels):
    ui, ul, rg, _ = swapped_info_batch(np.asarray(inputs).reshape(1, 128), np.asarray(labels).reshape(1, 128))
    return ui[0], ul[0], rg[0]


def miracle_view(updated_inputs, updated_labels, reduced_G, counter_stat):
    code = GL.get_map("code_parameters")
    hard = np.where(np.asarray(updated_inputs) > 0, 0, 1)
    mrb_error_num = int(((hard[:code.k] + np.asarray(updated_labels)[:code.k]) % 2).sum())
    counter_stat.update([mrb_error_num])
    return counter_stat, mrb_error_num


def _frames_of(selected_ds):
    """First row of every batch = the channel LLR of one failed frame (pb_testing.py:69-72)."""
    it = selected_ds.as_numpy_iterator() if hasattr(selected_ds, "as_numpy_iterator") else iter(selected_ds)
    ys, labs = [], []
    for batch in it:
        ys.append(np.asarray(batch[0])[0])
        labs.append(np.asarray(batch[1])[0])
    return np.asarray(ys, dtype=np.float32).reshape(-1, 128), np.asarray(labs).reshape(-1, 128)


def pb_osd_batch(inputs, labels, snr, order_limit):
    """PB-OSD on [B,128] channel LLRs -> dict(correct, num_teps, suc1, suc2, list_cmp, codeword)."""
    y = np.ascontiguousarray(np.asarray(inputs, dtype=np.float32).reshape(-1, 128))
    B = y.shape[0]
    h = get_handle()
    cw = np.empty((B, 4), np.uint32)
    stats = np.empty((B, 4), np.int32)
    h.call("ldpcb_osd_pb_decode_host", y, B, int(order_limit), float(snr), cw, stats)
    codeword = _lib.unpack_bits(cw)
    correct = (codeword == (np.asarray(labels).reshape(B, 128) & 1)).all(axis=1)
    return {"correct": correct, "num_teps": stats[:, 0], "suc1": stats[:, 1], "suc2": stats[:, 2], "list_cmp": stats[:, 3],
            "codeword": codeword}


def pb_osd(snr, selected_ds):
    start_time = time.process_time()
    order_limit = GL.get_map("order_limit")
    y, lab = _frames_of(selected_ds)
    logdir = "./log/"
    os.makedirs(logdir, exist_ok=True)
    summary = {"snr": snr, "order_limit": order_limit, "frames": len(y)}
    if GL.get_map("miracle_view"):
        ui, ul, rg, _ = swapped_info_batch(y, lab)
        hard = np.where(ui > 0, 0, 1)
        errs = ((hard[:, :64] + ul[:, :64]) % 2).sum(axis=1)
        counter_stat = Counter(int(e) for e in errs)
        total, acc = sum(counter_stat.values()), 0
        print("\nFor miracle view %.1fdB (order_limit:%d) :" % (snr, order_limit))
        print(f"total_sum:{total}")
        for key, value in sorted(counter_stat.items()):
            acc += value
            print(f"order-{key}: Accumulated Ratio: {acc/total:.4f}")
        summary["miracle"] = dict(counter_stat)
        return summary
    limit = GL.get_map("termination_num_threshlod") or 100
    if GL.get_map("pb_osd") and not GL.get_map("convention_osd"):
        res = pb_osd_batch(y, lab, snr, order_limit)
        fails_cum = np.cumsum(~res["correct"])
        n_used = int(np.searchsorted(fails_cum, limit) + 1) if fails_cum.size and fails_cum[-1] >= limit else len(y)
        ok = res["correct"][:n_used]
        correct_sum, fail_sum = int(ok.sum()), int((~ok).sum())
        actual = max(correct_sum + fail_sum, 1)
        FER = round(fail_sum / actual, 4)
        average_size = round(float(res["num_teps"][:n_used].sum()) / actual, 5)
        average_num_memory = round(float(res["list_cmp"][:n_used].sum()) / actual, 5)
        a1 = round(float(res["suc1"][:n_used].sum()) / actual, 5)
        a2 = round(float(res["suc2"][:n_used].sum()) / actual, 5)
        T2 = time.process_time()
        log_filename = logdir + "PB-OSD-order-" + str(order_limit) + ".txt"
        print("\nFor PB-OSD %.1fdB (order_limit:%d) :\n" % (snr, order_limit))
        print("----> S:" + str(correct_sum) + " F:" + str(fail_sum) + "\n")
        print(f"FER:{FER:.4f} Average TEPs:{average_size:.2f} Maintained_list_len:{average_num_memory:.2f} Average_suc: {a1:.2f}/{a2:.2f}")
        with open(log_filename, "a+") as f:
            f.write("\nFor PB-OSD %.1fdB (order_limit:%d) summary:\n" % (snr, order_limit))
            f.write(f"--> S/F:{correct_sum}/{fail_sum}\n")
            f.write(f"FER:{FER:.5f} Average TEPs:{average_size:.2f} Maintained_list_len:{average_num_memory:.2f} Average_suc: {a1:.2f}/{a2:2f}\n")
            f.write(f"Running time:{T2 - start_time} seconds with mean time {(T2 - start_time)/actual:.4f}!\n")
        summary.update({"S": correct_sum, "F": fail_sum, "FER": FER, "average_teps": average_size, "log": log_filename})
        return summary
    res = cnv_OSD.convention_osd_batch(y, lab, order_limit)
    # the reference stops after `limit` OSD failures (pb_testing.py:174, PB_OSD/globalmap.py:43)
    fails_cum = np.cumsum(~res["correct"])
    n_used = int(np.searchsorted(fails_cum, limit) + 1) if fails_cum.size and fails_cum[-1] >= limit else len(y)
    ok = res["correct"][:n_used]
    S, F = int(ok.sum()), int((~ok).sum())
    counter = Counter(int(p) for p in res["phase"][:n_used])
    FER = round(F / max(S + F, 1), 4)
    T2 = time.process_time()
    tag = "CNV-OSD" if GL.get_map("convention_osd") else "OSD(exhaustive sweep: neither pb_osd nor convention_osd is set)"
    log_filename = logdir + ("CNV-OSD-order-" if GL.get_map("convention_osd") else "PB-OSD-order-") + str(order_limit) + ".txt"
    print("\nFor %s %.1fdB (order_limit:%d) :\n" % (tag, snr, order_limit))
    print("----> S:" + str(S) + " F:" + str(F) + "\n")
    print("Distribution of phases:" + str(counter) + "\n")
    print("FER:" + str(FER) + " Average TEPs size:", res["teps_size"], "\n")
    with open(log_filename, "a+") as f:
        f.write("\nFor %s %.1fdB (order_limit:%d) summary:\n" % (tag, snr, order_limit))
        f.write("----> S:" + str(S) + " F:" + str(F) + "\n")
        f.write("Distribution of phases:" + str(counter) + "\n")
        f.write("FER:" + str(FER) + " Average TEPs size:" + str(res["teps_size"]) + "\n")
        f.write(f"Running time:{T2 - start_time} seconds with mean time {(T2 - start_time)/max(S + F, 1):.4f}!\n")
    summary.update({"S": S, "F": F, "FER": FER, "phases": dict(counter), "teps_size": res["teps_size"], "log": log_filename})
    return summary
