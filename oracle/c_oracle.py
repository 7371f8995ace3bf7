"""ctypes wrapper of the C oracle (TEST INFRASTRUCTURE ONLY; see oracle/c/ldpc_oracle.c).

Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference leg import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libldpc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "ldpc_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        r = subprocess.run(["make", "-C", os.path.join(HERE, "c"), "-B"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building the C oracle failed:\n" + r.stdout + r.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def nms(y, H, iters=12, alpha=0.66943514, w_vc=1.0, w_marg=1.0, early_stop=False, traj=False, threads=0):
    y = np.ascontiguousarray(y, dtype=np.float32)
    B = y.shape[0]
    H8 = np.ascontiguousarray(np.asarray(H) & 1, dtype=np.uint8)
    hard = np.empty((B, 128), np.uint8)
    syn = np.empty(B, np.uint8)
    it = np.empty(B, np.uint8)
    tr = np.empty((B, iters + 1, 128), np.float32) if traj else None
    lib().oracle_nms(_p(y), C.c_int64(B), _p(H8), C.c_int(iters), C.c_float(alpha), C.c_float(w_vc), C.c_float(w_marg),
                     C.c_int(int(early_stop)), _p(hard), _p(syn), _p(it), _p(tr), C.c_int(threads))
    return {"hard": hard, "syndrome_nz": syn.astype(bool), "iters_used": it, "traj": tr}


def osd(yo, ys, G, teps_packed, flags=0, block_start=None, truth=None, threads=0, want_perm=True):
    yo = np.ascontiguousarray(yo, dtype=np.float32)
    ys = yo if ys is None else np.ascontiguousarray(ys, dtype=np.float32)
    B = yo.shape[0]
    G8 = np.ascontiguousarray(np.asarray(G) & 1, dtype=np.uint8)
    teps = np.ascontiguousarray(teps_packed, dtype=np.uint32)
    cw = np.empty((B, 128), np.uint8)
    bt = np.empty(B, np.int32)
    bq = np.empty(B, np.int64)
    ex = np.empty(B, np.int32)
    pm = np.empty((B, 128), np.uint8) if want_perm else None
    rg = np.empty((B, 64), np.uint64) if want_perm else None
    nb = 0 if block_start is None else len(block_start) - 1
    bs = None if block_start is None else np.ascontiguousarray(block_start, dtype=np.int32)
    bm = np.empty((B, nb), np.int64) if nb else None
    ba = np.empty((B, nb), np.int32) if nb else None
    tr = None if truth is None else np.ascontiguousarray(truth, dtype=np.uint8)
    tq = np.empty(B, np.int64) if truth is not None else None
    lib().oracle_osd(_p(yo), _p(ys), C.c_int64(B), _p(G8), _p(teps), C.c_int(len(teps)), _p(bs), C.c_int(nb), C.c_int(flags),
                     _p(cw), _p(bt), _p(bq), _p(ex), _p(pm), _p(rg), _p(bm), _p(ba), _p(tr), _p(tq), C.c_int(threads))
    return {"codeword": cw, "best_tep": bt, "best_score_q": bq, "score_exp": ex, "perm": pm, "redG": rg,
            "block_min_q": bm, "block_arg": ba, "truth_score_q": tq}
