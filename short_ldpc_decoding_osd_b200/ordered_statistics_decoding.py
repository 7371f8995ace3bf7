"""Drop-in for ``LDPC_128/DL_OSD_Testing_serial/ordered_statistics_decoding.py`` (class ``osd``), the OSD of
the paper's DL scheme: ordering by the DIA reliability, Gaussian elimination, TEP blocks ("order patterns")
along a decoding path, sliding-window early termination.

What runs where:

* GPU (libldpc_b200.so, ldpcb_osd_block_minima): ascending reliability order with tf.argsort's tie rule, the
  basis, re-encoding of every TEP of every block, exact discrepancies against the CHANNEL LLR, the minimum of
  every block (``acquire_min``, ``:153-162``) and the discrepancy of the transmitted codeword (``:181-184``),
  for all frames of the batch in one call.  The reference eliminates H in ascending order and takes the
  non-pivot positions as MRB (``:43-80``); the kernel eliminates G in the reversed order.  By matroid duality
  both give the same MRB set, and the DL outputs (success, windows, complexity) do not depend on the order of
  the LRB, which is the only thing that differs (tests/test_oracle_golden.py checks it against the
  reference's own H-based code).
* The window bookkeeping of ``sliding_osd`` (``:164-220``) over the <= 30 block minima of a frame and the
  6 -> 6 -> 2 window classifier: on the GPU (ldpcb_dl_window_policy, one thread per frame) when ``fcn`` is a
  ``nn_net.Predict_outlier_light`` (it exposes its two kernels), on the host for an arbitrary callable.  Success is ``global_min == discrepancy_sum_truth`` evaluated on exact integers; the
  reference's fp32 equality can report false failures (see the golden test).
"""
from __future__ import annotations

from itertools import chain, combinations
from itertools import product as _product
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from . import globalmap as GL
from .fill_matrix_info import gf2_systematic_form
from .runtime import get_handle

FLAGS_DL = _lib.OSD_TIES_HIGH_INDEX_FIRST | _lib.OSD_DISC_HARD_FROM_SCORE


def pack_dl_teps(error_patterns: np.ndarray) -> np.ndarray:
    """int[T,64] patterns over DL MRB indices (0 = LEAST reliable) -> packed TEP words over kernel MRB
    positions (0 = MOST reliable): index i maps to 63 - i."""
    m = np.asarray(error_patterns)
    out = np.full(m.shape[0], 0xFFFFFFFF, dtype=np.uint32)
    if (m.sum(axis=1) > 4).any():
        raise ValueError("order patterns of total weight > 4 are not supported")
    for r in range(m.shape[0]):
        pos = sorted(63 - int(c) for c in np.flatnonzero(m[r]))
        v = 0xFFFFFFFF
        for i, p in enumerate(pos):
            v = (v & ~(0xFF << (8 * i))) | (p << (8 * i))
        out[r] = v
    return out


class osd:
    def __init__(self, code):
        self.original_H = code.H
        self.n_dims = code.check_matrix_column
        self.k = code.k
        self.m = self.n_dims - self.k
        self.code = code

    # ordered_statistics_decoding.py:25-28
    def mag_input_gen(self, inputs):
        a = np.abs(np.asarray(inputs, dtype=np.float32))
        return np.argsort(a, axis=-1, kind="stable").astype(np.int32)

    # ordered_statistics_decoding.py:81-98
    def error_pattern_gen(self, direction, range_list):
        code = GL.get_map("code_parameters") or self.code
        iters = [list(combinations(range_list[i], v)) if v else [(-1,)] for i, v in enumerate(direction)]
        joined = list(_product(*iters))
        patterns = np.zeros((len(joined), code.k), dtype=int)
        for i, seq in enumerate(joined):
            idx = [x for x in chain.from_iterable(seq) if x != -1]
            patterns[i, idx] = 1
        return patterns

    # ordered_statistics_decoding.py:43-80, literal H-based elimination on the host (diagnostics only: the
    # decoding path below does not need it)
    def identify_mrb(self, order_H_list):
        code = GL.get_map("code_parameters") or self.code
        threshold_sum = GL.get_map("threshold_sum") or 3
        n, k = code.check_matrix_column, code.k
        idx_list, M_list, swap_len, swap_pos = [], [], [], []
        for H in np.asarray(order_H_list):
            R, record = gf2_systematic_form(np.copy(H))
            index_order = np.arange(n)
            for a, b in record:
                index_order[a], index_order[b] = index_order[b], index_order[a]
            MRB, LRB = index_order[-k:], index_order[:k]
            mrb_swap = np.argsort(MRB, kind="stable")
            idx_list.append(np.concatenate([index_order[:n - k], np.sort(MRB)]))
            M_list.append(R[:, -k:][:, mrb_swap])
            swap_len.append(int(np.where(MRB >= n - k, 0, 1).sum()))
            swap_pos.append(np.where(LRB >= (n - k) - 4 * threshold_sum, 1, 0))
        return idx_list, M_list, swap_len, swap_pos

    # ordered_statistics_decoding.py:141-151
    def sliding_window_ops(self, fcn, window, global_min, k):
        sorted_window = np.sort(np.asarray(window, dtype=np.float32).reshape(1, -1))
        expanded = np.append(sorted_window, np.float32(k)).reshape(1, -1).astype(np.float32)
        output_prb = np.asarray(fcn(expanded)).reshape(-1)
        early = bool(output_prb[1] > GL.get_map("soft_margin"))
        return early, min(global_min, float(np.min(window)))

    def block_minima(self, input_list, inputs, labels, tep_info):
        """GPU part: -> (block_min_q int64[B,nb], truth_q int64[B], score_exp int32[B], perm uint8[B,128])."""
        teps_list, acc_block_size = tep_info
        list_length = (GL.get_map("num_iterations") or 12) + 1
        order_metric = np.ascontiguousarray(np.asarray(inputs, dtype=np.float32).reshape(-1, 128))
        B = order_metric.shape[0]
        channel = np.ascontiguousarray(np.asarray(input_list, dtype=np.float32).reshape(B, list_length, 128)[:, 0, :])
        packed = np.concatenate([pack_dl_teps(b) for b in teps_list])
        starts = np.asarray(acc_block_size, dtype=np.int32)
        nb = len(teps_list)
        import torch  # device memory carrier only

        dev = f"cuda:{get_handle().device}"
        h = get_handle()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        bm = torch.empty((B, nb), dtype=torch.int64, device=dev)
        ex = torch.empty((B,), dtype=torch.int32, device=dev)
        ts = torch.empty((B,), dtype=torch.int64, device=dev)
        pm = torch.empty((B, 128), dtype=torch.uint8, device=dev)
        truth = t(_lib.pack_bits(np.asarray(labels).reshape(B, 128)).view(np.int32))
        maxw = min(4, max(1, max(int(np.asarray(b).sum(axis=1).max()) for b in teps_list)))  # hint: rows touched per TEP
        h.call("ldpcb_osd_block_minima", t(order_metric), t(channel), B, t(packed.view(np.int32)), len(packed), t(starts), nb,
               FLAGS_DL | (maxw << _lib.OSD_MAXW_SHIFT), bm, None, ex, truth, ts, pm, None)
        torch.cuda.synchronize()
        return bm.cpu().numpy(), ts.cpu().numpy(), ex.cpu().numpy(), pm.cpu().numpy()

    def sliding_osd_gpu(self, fcn, input_list, inputs, labels, tep_info):
        """sliding_osd with the window policy on the GPU too (ldpcb_dl_window_policy); `fcn` must expose the two
        Keras kernels as W1 [(w+1),(w+1)] and W2 [(w+1),2] (nn_net.Predict_outlier_light)."""
        import torch  # device memory carrier only

        teps_list, acc_block_size = tep_info
        width = GL.get_map("sliding_win_width")
        bm, truth_q, ex, _ = self.block_minima(input_list, inputs, labels, tep_info)
        h = get_handle()
        dev = f"cuda:{h.device}"
        B, nb = bm.shape
        cnt = torch.zeros(4, dtype=torch.int64, device=dev)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        h.call("ldpcb_dl_window_policy", t(bm), t(ex), t(truth_q), B, nb, int(width),
               np.ascontiguousarray(fcn.W1, dtype=np.float32), np.ascontiguousarray(fcn.W2, dtype=np.float32),
               float(GL.get_map("soft_margin")), np.ascontiguousarray(acc_block_size, dtype=np.int32), None, None, None, cnt, None)
        torch.cuda.synchronize()
        c = cnt.cpu().numpy()
        return int(c[0]), int(c[1]), int(c[2]), int(c[3])

    # ordered_statistics_decoding.py:164-220
    def sliding_osd(self, fcn, input_list, inputs, labels, tep_info):
        if hasattr(fcn, "W1") and hasattr(fcn, "W2"):
            return self.sliding_osd_gpu(fcn, input_list, inputs, labels, tep_info)
        teps_list, acc_block_size = tep_info
        width = GL.get_map("sliding_win_width")
        bm, truth_q, ex, _ = self.block_minima(input_list, inputs, labels, tep_info)
        success_dec = failure_dec = 0
        complexity_sum = windows_sum = 0
        nblk = len(teps_list)
        for i in range(bm.shape[0]):
            scale = 2.0 ** (int(ex[i]) - 54)
            mins_q = [int(v) for v in bm[i]]
            window_q = mins_q[:width]
            global_q = min(window_q)
            deep_limit = width
            for k in range(nblk - width + 1):
                deep_limit = k + width
                if k != 0:
                    ms = mins_q[width + k - 1]
                    window_q = (window_q + [ms])[-width:]
                    if ms > global_q:
                        continue
                window = [np.float32(v * scale) for v in window_q]
                early, _ = self.sliding_window_ops(fcn, window, global_q * scale, k)
                global_q = min(global_q, min(window_q))
                if early:
                    break
            windows_sum += deep_limit - width + 1
            complexity_sum += int(acc_block_size[deep_limit])
            if global_q == int(truth_q[i]):
                success_dec += 1
            else:
                failure_dec += 1
        return success_dec, failure_dec, windows_sum, complexity_sum
