"""CPU oracle (TEST INFRASTRUCTURE ONLY) for ordered-statistics decoding.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product (short_ldpc_decoding_osd_b200/) never does.

NumPy/Python restatement of the reference's per-frame OSD:
  swapped_info / identify_mrb / full_gf2elim   LDPC_128/PB_OSD/pb_testing.py:231-320
  generate_teps / convention_osd_main          LDPC_128/FS_OSD/convention_osd.py:13-77
  generate_sequential_teps                     LDPC_128/FS_OSD/fs_testing.py:32-49
  osd.error_pattern_gen / acquire_min          LDPC_128/DL_OSD_Testing_serial/ordered_statistics_decoding.py:81-98,153-162

Parity status: the reference ships no tests or golden vectors.  This
restatement is pinned by tests/golden/osd_ref_shim.npz (the reference's own pb_testing.py /
convention_osd.py source run under a NumPy emulation of TensorFlow, oracle/ref_runner.py) and by
tests/golden/gf2elim_ref.npz (the reference's full_gf2elim lifted with `ast` and run unmodified).

One deliberate definition (DESIGN.md "exact score"): TensorFlow's fp32 reduce_sum has no specified
summation order, so "the TEP the reference chooses" is only defined up to fp32 rounding.  The oracle
scores with exact integers instead: every |y| is multiplied by 2^(54-E) (E = frexp exponent of the
frame's largest |y|) and rounded to nearest-even, sums are exact int64, and argmin takes the first
minimum in enumeration order (tf.argmin).  This equals the real-number argmin except when two
candidates differ by less than 2^-49 of the largest |y|.
"""
from __future__ import annotations

from itertools import combinations, product
from typing import List, Sequence, Tuple

import numpy as np

N, K = 128, 64
TIES_HIGH_INDEX_FIRST = 1
DISC_HARD_FROM_SCORE = 2


# ---- reliability sort -------------------------------------------------------------------------------
def reliability_order(y: np.ndarray, ties_high_index_first: bool = False) -> np.ndarray:
    """tf.argsort(|y|, DESCENDING) (pb_testing.py:308-310): stable, lower index first on ties.

    With ties_high_index_first the result is the exact reverse of tf.argsort(|y|, ASCENDING)
    (ordered_statistics_decoding.py:25-28).  Keys are the raw bits of |y| (monotone for finite fp32).
    """
    key = np.abs(np.asarray(y, dtype=np.float32)).view(np.uint32).astype(np.int64)
    if ties_high_index_first:
        asc = np.argsort(key, kind="stable")
        return asc[::-1].copy()
    return np.argsort(-key, kind="stable")


# ---- GF(2) elimination, the reference's rule restated -----------------------------------------------
def full_gf2elim(M: np.ndarray) -> Tuple[np.ndarray, List[Tuple[int, int]]]:
    """pb_testing.py:231-266 restated (row swap with the first 1 at/below the diagonal, else column
    swap with the first 1 of row i, then XOR row i into every other row with a 1 in column j)."""
    A = np.array(M, dtype=np.int64) & 1
    m, n = A.shape
    i = j = 0
    swaps: List[Tuple[int, int]] = []
    while i < m and j < n:
        colpart = A[i:, j]
        if colpart.max():
            k = int(np.argmax(colpart)) + i
            if k != i:
                A[[i, k]] = A[[k, i]]
        else:
            rowpart = A[i, j:]
            if not rowpart.max():
                A = np.delete(A, i, axis=0)
                m -= 1
                continue
            ck = int(np.argmax(rowpart)) + j
            A[:, [j, ck]] = A[:, [ck, j]]
            swaps.append((j, ck))
        hit = A[:, j].copy()
        hit[i] = 0
        A[hit == 1, j:] ^= A[i, j:]
        i += 1
        j += 1
    return A, swaps


def identify_mrb(order_G: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """pb_testing.py:268-304 -> (updated_G int[64,128] = [I | P'], updated_index_order int[128])."""
    swapped_G, swaps = full_gf2elim(order_G)
    index_order = np.arange(N)
    for a, b in swaps:
        index_order[a], index_order[b] = index_order[b], index_order[a]
    mrb = index_order[:K]
    mrb_swapping = np.argsort(mrb, kind="stable")
    mrb_order = np.sort(mrb)
    ident = np.identity(K, dtype=np.int64)
    updated_mrb_matrix = ident[:, mrb_swapping]
    lrb = index_order[K:]
    lrb_swapping = np.argsort(lrb, kind="stable")
    lrb_order = np.sort(lrb)
    interm = swapped_G[:, K:][:, lrb_swapping]
    updated_lrb = (updated_mrb_matrix.T.dot(interm)) % 2
    updated_G = np.concatenate([ident, updated_lrb], axis=1)
    return updated_G, np.concatenate([mrb_order, lrb_order])


def swapped_info(y: np.ndarray, labels: np.ndarray, G: np.ndarray, ties_high_index_first: bool = False):
    """pb_testing.py:306-320 -> (updated_inputs f32[128], updated_labels, reduced_G, perm int[128])."""
    y = np.asarray(y, dtype=np.float32)
    pi1 = reliability_order(y, ties_high_index_first)
    order_G = np.asarray(G)[:, pi1]
    reduced_G, pi2 = identify_mrb(order_G)
    perm = pi1[pi2]
    return y[perm], np.asarray(labels)[perm], reduced_G, perm


def greedy_mrb(y: np.ndarray, G: np.ndarray, ties_high_index_first: bool = False):
    """Textbook greedy most-reliable basis; used to cross-check identify_mrb (same set)."""
    pi1 = reliability_order(y, ties_high_index_first)
    A = (np.asarray(G)[:, pi1] & 1).astype(np.uint8)
    used = np.zeros(K, dtype=bool)
    piv = []
    for c in range(N):
        rows = np.flatnonzero(A[:, c] & ~used)
        if rows.size == 0:
            continue
        p = rows[0]
        used[p] = True
        piv.append(c)
        hit = A[:, c].astype(bool)
        hit[p] = False
        A[hit] ^= A[p]
        if len(piv) == K:
            break
    return pi1, np.array(piv)


# ---- TEP enumerations ------------------------------------------------------------------------------------
def generate_teps_conv(order: int, k: int = K) -> List[Tuple[int, ...]]:
    """convention_osd.py:13-38: per weight, combinations sorted by descending index sum (stable)."""
    out: List[Tuple[int, ...]] = []
    for w in range(order + 1):
        cs = list(combinations(range(k), w))
        cs.sort(key=lambda c: -sum(c))
        out.extend(cs)
    return out


def generate_teps_fs(order: int, k: int = K) -> List[Tuple[int, ...]]:
    """All-zero TEP (fs_testing.py:131-132) then generate_sequential_teps (fs_testing.py:32-49):
    lexicographic combinations with the vector reversed, i.e. support {c} -> {k-1-c}."""
    out: List[Tuple[int, ...]] = [()]
    for w in range(1, order + 1):
        for c in combinations(range(k), w):
            out.append(tuple(sorted(k - 1 - x for x in c)))
    return out


def boundary_list(order: int, k: int = K) -> List[int]:
    """query_boundary (convention_osd.py:39-47): [1, 65, 2081, 43745][:order+1]."""
    from math import comb

    acc, out = 0, []
    for w in range(order + 1):
        acc += comb(k, w)
        out.append(acc)
    return out


def dl_segments(k: int = K, num_seg: int = 6):
    """secure_segment_threshold (DL_OSD_Testing_serial/globalmap.py:57-76) -> sizes, boundaries."""
    allocation = k - 1
    basic = list(range(1, num_seg))
    nb = sum(basic)
    sizes = [int(allocation / nb * b) for b in basic]
    sizes[-1] += allocation - sum(sizes)
    sizes = [1] + sizes
    return sizes, [0] + list(np.cumsum(sizes))


def dl_error_pattern_block(direction: Sequence[int], range_list) -> List[Tuple[int, ...]]:
    """osd.error_pattern_gen (ordered_statistics_decoding.py:81-98): Cartesian product, in
    itertools.product order, of combinations(segment_i, w_i); indices are DL MRB positions
    (0 = LEAST reliable)."""
    iters = [list(combinations(range_list[i], v)) if v else [()] for i, v in enumerate(direction)]
    return [tuple(x for part in combo for x in part) for combo in product(*iters)]


def pack_teps(teps: Sequence[Sequence[int]], dl_index: bool = False) -> np.ndarray:
    """Packed uint32 TEP words of include/ldpc_b200.h (byte i = i-th position ascending, 0xFF unused).
    dl_index: positions are DL indices (0 = least reliable) and are mapped to 63-i."""
    out = np.full(len(teps), 0xFFFFFFFF, dtype=np.uint32)
    for n, t in enumerate(teps):
        pos = sorted((K - 1 - x) if dl_index else x for x in t)
        v = 0xFFFFFFFF
        for i, p in enumerate(pos):
            v = (v & ~(0xFF << (8 * i))) | (p << (8 * i))
        out[n] = v
    return out


# ---- exact score ---------------------------------------------------------------------------------------
def score_abs(y: np.ndarray) -> np.ndarray:
    a = np.abs(np.asarray(y, dtype=np.float32))
    a = np.where(np.isnan(a), np.float32(0), a)
    return np.minimum(a, np.float32(3.402823466e38)).astype(np.float32)


def quantize(y_score: np.ndarray) -> Tuple[np.ndarray, int]:
    """|y| -> exact int64 weights q = rint(|y| * 2^(54-E)), E = frexp exponent of max |y|."""
    a = score_abs(y_score)
    amax = float(a.max()) if a.size else 0.0
    E = int(np.frexp(np.float32(amax))[1])
    q = np.rint(np.ldexp(a.astype(np.float64), 54 - E)).astype(np.int64)
    return q, E


def hard_of(y: np.ndarray) -> np.ndarray:
    """tf.where(y > 0, 0, 1) (convention_osd.py:54)."""
    return np.where(np.asarray(y, dtype=np.float32) > 0, 0, 1).astype(np.int64)


def osd_frame(y_order, y_score, G, teps: Sequence[Sequence[int]], flags: int = 0, block_start=None, truth=None):
    """One frame of exhaustive OSD over `teps` (MRB positions, 0 = most reliable).

    Returns dict: perm[128], reduced_G[64,128], best_tep, best_score_q, score_exp, codeword[128]
    (original positions), and with block_start: block_min_q[], block_arg[]; with truth: truth_score_q.
    Follows convention_osd_main (FS_OSD/convention_osd.py:49-77) with the exact score.
    """
    y_order = np.asarray(y_order, dtype=np.float32)
    y_score = np.asarray(y_score, dtype=np.float32)
    ties_high = bool(flags & TIES_HIGH_INDEX_FIRST)
    _, _, reduced_G, perm = swapped_info(y_order, np.zeros(N, dtype=np.int64), G, ties_high)
    yo, ys = y_order[perm], y_score[perm]
    q, E = quantize(ys)
    ho = hard_of(yo)
    hd = hard_of(ys) if (flags & DISC_HARD_FROM_SCORE) else ho
    # bit-packed rows of reduced_G as Python ints (bit t = permuted position t)
    rows = [int("".join(str(int(b)) for b in reduced_G[t][::-1]), 2) for t in range(K)]
    hd_int = int("".join(str(int(b)) for b in hd[::-1]), 2)
    c0 = 0
    for t in range(K):
        if ho[t]:
            c0 ^= rows[t]
    qs = [int(v) for v in q]

    def score_of(cw: int) -> int:
        d = cw ^ hd_int
        s = 0
        while d:
            low = d & -d
            s += qs[low.bit_length() - 1]
            d ^= low
        return s

    scores = np.empty(len(teps), dtype=np.int64)
    for n, t in enumerate(teps):
        cw = c0
        for p in t:
            cw ^= rows[p]
        scores[n] = score_of(cw)
    res = {"perm": perm.astype(np.uint8), "reduced_G": reduced_G, "score_exp": E}
    best = int(np.argmin(scores)) if len(teps) else -1
    res["best_tep"] = best
    res["best_score_q"] = int(scores[best]) if best >= 0 else None
    cw = c0
    for p in (teps[best] if best >= 0 else ()):
        cw ^= rows[p]
    cw_perm = np.array([(cw >> t) & 1 for t in range(N)], dtype=np.uint8)
    codeword = np.zeros(N, dtype=np.uint8)
    codeword[perm] = cw_perm
    res["codeword"] = codeword
    res["scores"] = scores
    if block_start is not None:
        bm, ba = [], []
        for b in range(len(block_start) - 1):
            seg = scores[block_start[b]:block_start[b + 1]]
            a = int(np.argmin(seg))
            bm.append(int(seg[a]))
            ba.append(a + int(block_start[b]))
        res["block_min_q"] = np.array(bm, dtype=np.int64)
        res["block_arg"] = np.array(ba, dtype=np.int32)
    if truth is not None:
        tperm = np.asarray(truth)[perm].astype(np.int64)
        res["truth_score_q"] = int(np.sum(q[(tperm ^ hd) == 1]))
    return res


def convention_osd_main(updated_inputs, updated_labels, reduced_G, teps, boundaries):
    """FS_OSD/convention_osd.py:49-77 on already permuted data, exact score
    -> (correct_indicator, teps_size, belonged_phase, estimated_index)."""
    yo = np.asarray(updated_inputs, dtype=np.float32)
    q, _ = quantize(yo)
    h = hard_of(yo)
    Gr = np.asarray(reduced_G).astype(np.int64)
    tepm = np.zeros((len(teps), K), dtype=np.int64)
    for n, t in enumerate(teps):
        tepm[n, list(t)] = 1
    mrb = (tepm + h[None, :K]) % 2
    cands = mrb.dot(Gr) % 2
    disc = (cands + h[None]) % 2
    scores = disc.dot(q)
    idx = int(np.argmin(scores))
    ok = bool(np.all(cands[idx] == np.asarray(updated_labels).astype(np.int64)))
    phase = -1
    if ok:
        for i, b in enumerate(boundaries):
            if idx < b:
                phase = i
                break
    return ok, len(teps), phase, idx


# ---- FS-OSD policy -----------------------------------------------------------------------------------------
def fs_frame(y, G, order_limit: int, tau_e: float = 6.5, tau_psc: int = 30, beta: float = 0.1, labels=None):
    """One frame of fs_osd (FS_OSD/fs_testing.py:94-165) with the exact integer score.

    Returns dict(codeword[128] original positions, best_tep (index in generate_teps_fs order), num_teps,
    stop_kind: 0 order-0 accepted / 1 tau_e stop / 2 skip rule / 3 all orders swept, success if labels given).
    The reference compares fp32 sums; here boundary, beta*(n-k) and w_dmin live in the frame's integer
    score units (beta*(n-k) is first rounded to fp32 as in `x + beta*(n-k)` with x an fp32 tensor).
    """
    y = np.asarray(y, dtype=np.float32)
    _, _, reduced_G, perm = swapped_info(y, np.zeros(N, dtype=np.int64), G)
    yp = y[perm]
    q, E = quantize(yp)
    h = hard_of(yp)
    rows = [int("".join(str(int(b)) for b in reduced_G[t][::-1]), 2) for t in range(K)]
    h_int = int("".join(str(int(b)) for b in h[::-1]), 2)
    qs = [int(v) for v in q]
    c0 = 0
    for t in range(K):
        if h[t]:
            c0 ^= rows[t]

    def evaluate(tep):
        cw = c0
        for p in tep:
            cw ^= rows[p]
        d = cw ^ h_int
        hd = bin(d).count("1")
        s = 0
        dd = d
        while dd:
            low = dd & -dd
            s += qs[low.bit_length() - 1]
            dd ^= low
        return cw, hd, s

    shift = float(np.float32(beta * (N - K)))
    sh = np.ldexp(np.float64(shift), 54 - E)
    shift_q = (1 << 62) if sh >= 4.6e18 else int(np.rint(sh))
    teps = generate_teps_fs(order_limit)
    cls = [0, 1, 65, 2081, 43745]
    opt_cw, hd0, w_dmin = evaluate(())
    best, num, kind = 0, 1, 3
    if hd0 < tau_e:
        kind = 0
    else:
        bnd = 0
        stop = False
        for j in range(order_limit):
            bnd += qs[63 - j]
            if not (bnd + shift_q < w_dmin):
                kind = 2
                break
            for i in range(cls[j + 1], cls[j + 2]):
                num += 1
                cw, hd, s = evaluate(teps[i])
                if hd < tau_e:
                    stop = True
                    kind = 1
                    break
                if hd < tau_psc and s < w_dmin:
                    w_dmin, opt_cw, best = s, cw, i
            if stop:
                break
    cw_perm = np.array([(opt_cw >> t) & 1 for t in range(N)], dtype=np.uint8)
    codeword = np.zeros(N, dtype=np.uint8)
    codeword[perm] = cw_perm
    out = {"codeword": codeword, "best_tep": best, "num_teps": num, "stop_kind": kind, "best_score_q": w_dmin, "score_exp": E,
           "perm": perm.astype(np.uint8)}
    if labels is not None:
        out["success"] = bool(np.array_equal(codeword, np.asarray(labels).astype(np.uint8)))
    return out
