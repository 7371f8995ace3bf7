"""Drop-in for the reference's conventional order-p OSD (``LDPC_128/FS_OSD/convention_osd.py`` and the
``PB_OSD`` copy, which differs only in taking a sixth tuple element ``updated_original_inputs`` whose
magnitudes weight the discrepancy, ``PB_OSD/convention_osd.py:50,61``).

Same functions and return values:

* ``generate_teps(order_limit)`` -> ``int32[T,64]`` TEP matrix in the reference's order (``:31-38``)
* ``query_boundary(order_limit)`` -> ``[1, 65, 2081, ...]`` (``:39-47``)
* ``convention_osd_main(wrapped_input)`` -> ``(correct_indicator, teps_size, belonged_phase)`` (``:49-77``);
  accepts the 5-tuple of the FS copy and the 6-tuple of the PB copy.

plus the batched form the drivers use, ``convention_osd_batch`` (all frames of a run in one library call).
The sweep, re-encoding, discrepancy and argmin run in libldpc_b200.so (ldpcb_osd_sweep_host /
ldpcb_osd_decode_host); scores are exact integers, so "first minimum" is well defined.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from . import globalmap as GL
from .runtime import get_handle

_TEP_CACHE = {}


def binomial_coefficient(n: int, k: int) -> int:
    return math.factorial(n) // (math.factorial(k) * math.factorial(n - k))


def unpack_tep_words(words: np.ndarray, k: int = 64) -> np.ndarray:
    """Packed TEP words (include/ldpc_b200.h) -> int32[T,k] 0/1 matrix."""
    out = np.zeros((len(words), k), dtype=np.int32)
    for j in range(4):
        pos = (words >> np.uint32(8 * j)) & np.uint32(0xFF)
        rows = np.flatnonzero(pos < k)
        out[rows, pos[rows]] = 1
    return out


def pack_tep_matrix(error_patterns_matrix) -> np.ndarray:
    """int[T,64] 0/1 TEP matrix -> packed words; weight > 4 is not supported by the sweep kernel."""
    m = np.asarray(error_patterns_matrix)
    if m.ndim != 2 or m.shape[1] != 64:
        raise ValueError(f"TEP matrix must be [T,64], got {m.shape}")
    if (m.sum(axis=1) > 4).any():
        raise ValueError("TEPs of weight > 4 are not supported")
    out = np.full(m.shape[0], 0xFFFFFFFF, dtype=np.uint32)
    rows, cols = np.nonzero(m)
    slot = np.zeros(m.shape[0], dtype=np.int64)
    for r, c in zip(rows, cols):  # np.nonzero is row-major, so positions arrive ascending per row
        s = slot[r]
        out[r] = (out[r] & ~np.uint32(0xFF << (8 * s))) | np.uint32(int(c) << (8 * s))
        slot[r] += 1
    return out


def generate_teps(order_limit: int) -> np.ndarray:
    key = ("conv", int(order_limit))
    if key not in _TEP_CACHE:
        h = get_handle()
        _TEP_CACHE[key] = unpack_tep_words(h.tep_table(int(order_limit), _lib.TEP_CONV))
    return _TEP_CACHE[key]


def query_boundary(order_limit: int) -> List[int]:
    code = GL.get_map("code_parameters")
    k = code.k if code is not None else 64
    out, acc = [], 0
    for i in range(order_limit + 1):
        acc += binomial_coefficient(k, i)
        out.append(acc)
    return out


def _redG_words(reduced_G) -> np.ndarray:
    g = np.asarray(reduced_G)
    if g.shape != (64, 128):
        raise ValueError(f"reduced_G must be [64,128], got {g.shape}")
    if not np.array_equal(g[:, :64] & 1, np.identity(64, dtype=g.dtype)):
        raise ValueError("reduced_G must be systematic: [I | P']")
    bits = np.ascontiguousarray(g[:, 64:] & 1, dtype=np.uint8)
    return np.packbits(bits, axis=1, bitorder="little").view("<u8").reshape(64)


def convention_osd_main(wrapped_input) -> Tuple[bool, int, int]:
    if len(wrapped_input) == 6:
        updated_inputs, updated_original_inputs, updated_labels, reduced_G, error_patterns_matrix, boundary_list = wrapped_input
    else:
        updated_inputs, updated_labels, reduced_G, error_patterns_matrix, boundary_list = wrapped_input
        updated_original_inputs = updated_inputs
    yo = np.ascontiguousarray(np.asarray(updated_inputs, dtype=np.float32).reshape(1, 128))
    ys = yo if updated_original_inputs is updated_inputs else np.ascontiguousarray(np.asarray(updated_original_inputs, dtype=np.float32).reshape(1, 128))
    words = np.ascontiguousarray(_redG_words(reduced_G).reshape(1, 64))
    teps = pack_tep_matrix(error_patterns_matrix)
    h = get_handle()
    cw = np.empty((1, 4), np.uint32)
    best = np.empty(1, np.int32)
    # hard decisions of the discrepancy come from updated_inputs in both copies (convention_osd.py:54,60)
    h.call("ldpcb_osd_sweep_host", yo, ys, words, 1, teps, len(teps), 0, cw, best, None, None)
    cand = _lib.unpack_bits(cw)[0]
    labels = np.asarray(updated_labels).reshape(-1).astype(np.int64)
    correct_indicator = bool(np.array_equal(cand.astype(np.int64), labels))
    belonged_phase = -1
    if correct_indicator:
        for i, b in enumerate(boundary_list):
            if int(best[0]) < b:
                belonged_phase = i
                break
    return correct_indicator, len(teps), belonged_phase


def convention_osd_batch(inputs, labels, order_limit: int, original_inputs=None, tep_order: int = _lib.TEP_CONV):
    """All frames at once: reliability sort + elimination + order-p sweep on the GPU.

    inputs float[B,128] (ordering metric; also the scoring metric unless original_inputs is given),
    labels int[B,128].  Returns dict(correct bool[B], best_tep int32[B], phase int[B], codeword uint8[B,128],
    perm uint8[B,128], teps_size int).
    """
    yo = np.ascontiguousarray(np.asarray(inputs, dtype=np.float32).reshape(-1, 128))
    ys = yo if original_inputs is None else np.ascontiguousarray(np.asarray(original_inputs, dtype=np.float32).reshape(-1, 128))
    B = yo.shape[0]
    h = get_handle()
    cw = np.empty((B, 4), np.uint32)
    best = np.empty(B, np.int32)
    perm = np.empty((B, 128), np.uint8)
    h.call("ldpcb_osd_decode_host", yo, ys, B, int(order_limit), int(tep_order), 0, cw, best, None, None, perm, None)
    codeword = _lib.unpack_bits(cw)
    correct = (codeword == (np.asarray(labels).reshape(B, 128) & 1)).all(axis=1)
    bnd = np.array(query_boundary(order_limit))
    phase = np.where(correct, np.searchsorted(bnd, best, side="right"), -1)
    return {"correct": correct, "best_tep": best, "phase": phase, "codeword": codeword, "perm": perm,
            "teps_size": h.tep_count(int(order_limit), int(tep_order))}
