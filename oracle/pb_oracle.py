"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the PB-OSD policy of LDPC_128/PB_OSD/pb_testing.py:100-149 with
optimal_tep_sequence (:366-397) and the probability terms (:35-41, :399-500).

Best-first TEP order and weighted distances use the exact integer reliabilities (oracle/osd_oracle.quantize);
the probabilities are evaluated in fp32 in the order the reference evaluates them, the binomial CDFs in fp64
(scipy in the reference).  Pinned by tests/golden/pb_ref_shim.npz (the reference's pb_osd under the TF shim).
"""
from __future__ import annotations

from math import comb

import numpy as np

from . import osd_oracle as OO

F32 = np.float32
N, K = 128, 64


def sigmoid32(x):
    x = np.asarray(x, dtype=F32)
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def binom_cdf_table(p: float, n: int = 64) -> np.ndarray:
    """cdf[b] = P(X <= b), X ~ Bin(n, p), fp64 (scipy.stats.binom.cdf in the reference)."""
    p = float(p)
    pmf = np.array([comb(n, i) * p ** i * (1.0 - p) ** (n - i) for i in range(n + 1)], dtype=np.float64)
    return np.minimum(np.cumsum(pmf), 1.0)


def pb_frame(y, G, snr: float, order_limit: int, labels=None):
    y = np.asarray(y, dtype=F32)
    _, _, reduced_G, perm = OO.swapped_info(y, np.zeros(N, dtype=np.int64), G)
    yp = y[perm]
    a = np.abs(yp).astype(F32)
    q, E = OO.quantize(yp)
    qs = [int(v) for v in q]
    h = OO.hard_of(yp)
    rows = [int("".join(str(int(b)) for b in reduced_G[t][::-1]), 2) for t in range(K)]
    h_int = int("".join(str(int(b)) for b in h[::-1]), 2)
    c0 = 0
    for t in range(K):
        if h[t]:
            c0 ^= rows[t]

    def score(cw):
        d = cw ^ h_int
        s, dd = 0, d
        while dd:
            low = dd & -dd
            s += qs[low.bit_length() - 1]
            dd ^= low
        return s, d

    nv = 1.0 / (10.0 ** (snr / 10.0))
    c4 = F32(-4.0 * nv)
    sig = sigmoid32(c4 * a)                                  # sigmoid(-4 nv |y_i|), fp32
    p1 = F32(np.mean(sig[K:], dtype=F32))                    # mean_lrb_prob (:346-351)
    pt = F32(np.mean(sig[:K], dtype=F32))                    # mean_mrb_prob (:449-454)
    niu = float(binom_cdf_table(float(pt), K)[order_limit])  # calculate_two_thresholds (:485-500)
    N_max = sum(comb(K, i) for i in range(order_limit + 1))
    p_t_suc = 0.99 * niu
    p_t_pro = 0.002 * np.sqrt((1.0 - niu) / N_max)
    spl = F32(1.0)
    for i in range(K):                                       # com_mrb_prob (:35-41)
        spl = F32(spl * F32(F32(1) - sig[i]))
    cdf_p1 = binom_cdf_table(float(p1), N - K)
    cdf_half = binom_cdf_table(0.5, N - K)
    lrb_mean = F32(np.mean(a[K:], dtype=F32))

    opt_cw = c0
    w_dmin, _ = score(c0)
    scale = 2.0 ** (E - 54)
    # list of live TEPs in insertion order: (exact MRB weight, tuple)
    lst = [(qs[K - 1], (K - 1,))]
    cost, early = 0, False
    suc1 = suc2 = list_cmp = 0
    for j in range(N_max - 1):
        list_cmp += 1 if len(lst) == 1 else 2
        m = min(range(len(lst)), key=lambda i: (lst[i][0], i))   # tf.argmin: first minimum
        wsum, tep = lst.pop(m)
        last, w = tep[-1], len(tep)
        if last < K - 1 and w < order_limit:
            lst.append((wsum + qs[K - 1], tep + (K - 1,)))
        if w > 1:
            if last - tep[-2] > 1:
                lst.append((wsum - qs[last] + qs[last - 1], tep[:-1] + (last - 1,)))
        elif last - 1 > -1:
            lst.append((qs[last - 1], (last - 1,)))
        cw = c0
        for p in tep:
            cw ^= rows[p]
        w_de, d = score(cw)
        # acquire_prob_promising (:431-447)
        rel = F32(0)
        for p in tep:
            rel = F32(rel + a[p])
        w1 = F32(np.exp(F32(c4 * rel), dtype=F32) * spl)
        w2 = F32(F32(1) - w1)
        tmp = np.floor(F32(F32(F32(w_dmin * scale) - rel) / lrb_mean))
        beta = int(min(max(0.0, float(tmp)), N - K))
        p_e_pro = float(F32(F32(F32(0.0) + F32(w1 * F32(cdf_p1[beta]))) + F32(w2 * F32(cdf_half[beta]))))
        if p_e_pro < p_t_pro:
            early, cost = True, j + 1
            break
        suc1 += 1
        if w_de < w_dmin:
            opt_cw, w_dmin = cw, w_de
            suc2 += 1
            # acquire_p_e_suc (:410-423)
            rel_d = F32(0)
            for t in range(K):
                if (d >> t) & 1:
                    rel_d = F32(rel_d + a[t])
            tep_p = F32(np.exp(F32(c4 * rel_d), dtype=F32) * spl)
            ratio = F32(F32(F32(1) - tep_p) / tep_p)
            prod = F32(1.0)
            for i in range(K, N):
                f = F32(F32(2) * (sig[i] if (d >> i) & 1 else F32(F32(1) - sig[i])))
                prod = F32(prod * f)
            p_e_suc = F32(F32(1) / F32(F32(1) + F32(ratio / prod)))
            if p_e_suc > F32(p_t_suc):
                early, cost = True, j + 1
                break
    num = cost if early else N_max
    cw_perm = np.array([(opt_cw >> t) & 1 for t in range(N)], dtype=np.uint8)
    codeword = np.zeros(N, dtype=np.uint8)
    codeword[perm] = cw_perm
    out = {"codeword": codeword, "num_teps": num, "early": early, "suc1": suc1, "suc2": suc2, "list_cmp": list_cmp,
           "best_score_q": w_dmin, "score_exp": E}
    if labels is not None:
        out["success"] = bool(np.array_equal(codeword, np.asarray(labels).astype(np.uint8)))
    return out
