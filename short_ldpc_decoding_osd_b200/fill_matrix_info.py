"""Code definition: alist parser, parity-check matrix H and systematic generator G.

Host-side mirror of the reference's ``Code`` class
(``LDPC_128/Ldpc_128_testing/fill_matrix_info.py:3-129``; eight near-identical copies
exist in the reference, one per script directory).  Same public attributes:
``H``, ``G``, ``k``, ``max_chk_degree``, ``check_matrix_row``, ``check_matrix_column``.

This runs once per process on the host (NumPy); it is not on the GPU hot path.  The
elimination rule is the reference's (``fill_matrix_info.py:7-42``): ``i == j`` march
together; a pivot is the *first* row at/below the diagonal holding a 1 (row swap);
if there is none, the column is swapped with the first later column holding a 1 in
row ``i`` and the swap is recorded; then the pivot row is XORed into every other row
holding a 1 in the pivot column (Gauss-Jordan).
"""
from __future__ import annotations

import hashlib
import os
from typing import List, Tuple

import numpy as np

__all__ = ["Code", "gf2_systematic_form", "CCSDS_ALIST"]

CCSDS_ALIST = os.path.join(os.path.dirname(__file__), "data", "CCSDS_ldpc_n128_k64.alist")


def gf2_systematic_form(M: np.ndarray) -> Tuple[np.ndarray, List[Tuple[int, int]]]:
    """Gauss-Jordan elimination over GF(2) with the reference's pivot rule.

    Returns the reduced matrix (all-zero rows removed) and the list of recorded
    column swaps ``(j, column_k)``.  Semantics follow
    ``fill_matrix_info.py:7-42`` / ``PB_OSD/pb_testing.py:231-266``.
    """
    A = np.array(M, dtype=np.uint8) & 1
    rows, cols = A.shape
    swaps: List[Tuple[int, int]] = []
    alive = list(range(rows))  # physical row ids still present, in logical order
    i = 0
    j = 0
    while i < len(alive) and j < cols:
        below = [r for r in alive[i:] if A[r, j]]
        if below:
            k = alive.index(below[0])
            alive[i], alive[k] = alive[k], alive[i]
        else:
            row = A[alive[i], j:]
            if not row.any():
                del alive[i]
                continue
            ck = int(np.flatnonzero(row)[0]) + j
            A[:, [j, ck]] = A[:, [ck, j]]
            swaps.append((j, ck))
        p = alive[i]
        hit = A[:, j].astype(bool)
        hit[p] = False
        A[hit] ^= A[p]
        i += 1
        j += 1
    return A[alive].astype(np.int64), swaps


class Code:
    """``Code(H_filename)`` -> H, G and Tanner-graph parameters (reference API)."""

    def __init__(self, H_filename: str = CCSDS_ALIST):
        self.load_code(H_filename)

    # reference name kept: fill_matrix_info.py:7
    def gf2elim(self, M):
        return gf2_systematic_form(M)

    # reference name kept: fill_matrix_info.py:44-69
    def generator_matrix(self, parity_check_matrix: np.ndarray) -> np.ndarray:
        R, swaps = gf2_systematic_form(parity_check_matrix)
        m, n = R.shape
        # R = [I | H2]  ->  G = [H2^T | I]  (fill_matrix_info.py:52-58)
        G = np.concatenate([R[:, m:].T, np.identity(n - m, dtype=np.int64)], axis=1)
        for a, b in reversed(swaps):  # undo column swaps (fill_matrix_info.py:59-64)
            G[:, [a, b]] = G[:, [b, a]]
        if np.any(parity_check_matrix.dot(G.T) % 2):
            raise ValueError("generator matrix failed the H.G^T = 0 check")
        return G

    # reference name kept: fill_matrix_info.py:70-129
    def load_code(self, H_filename: str) -> None:
        with open(H_filename, "rt") as f:
            tokens = [ln.split() for ln in f.read().splitlines()]
        n, m = int(tokens[0][0]), int(tokens[0][1])
        max_var_degree, max_chk_degree = int(tokens[1][0]), int(tokens[1][1])
        H = np.zeros((m, n), dtype=np.int64)
        # lines 4 .. 4+n-1: for each variable node the 1-based check indices (0 = padding)
        for v in range(n):
            for s in tokens[4 + v]:
                if s != "0":
                    H[int(s) - 1, v] = 1
        self.H = H
        self.max_chk_degree = max_chk_degree
        self.max_var_degree = max_var_degree
        self.check_matrix_column = n
        self.check_matrix_row = m
        self.G = self.generator_matrix(H)
        self.k = self.G.shape[0]

    # convenience (not in the reference)
    def sha256_prefixes(self) -> Tuple[str, str]:
        h = hashlib.sha256(self.H.astype(np.uint8).tobytes()).hexdigest()[:16]
        g = hashlib.sha256(self.G.astype(np.uint8).tobytes()).hexdigest()[:16]
        return h, g
