"""CPU check of the quasi-cyclic NMS kernel's data flow (csrc/nms_qc.cu): the class table, the lane relabelling
(rho, sigma), the rotation amounts and the ascending-check summation order are parsed from the source and checked
against the code's H -- every shuffle must connect a check's owner lane to the owner of a variable H says it touches,
each check must see its 8 variables and each variable its 3 or 5 checks exactly once, in ascending check order."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "short_ldpc_decoding_osd_b200", "csrc", "nms_qc.cu")).read()


def _table(name, dims):
    m = re.search(r"constexpr int T" + re.escape(dims) + r" = (\{.*?\});", SRC[SRC.index(name):], flags=re.S)
    return eval(m.group(1).replace("{", "[").replace("}", "]"))


EDGE = _table("constexpr CS edge(", "[4][8][2]")
RHO = _table("constexpr int rho(", "[4]")
SIG = _table("constexpr int sig(", "[8]")


def delta(R, e):
    C, s = EDGE[R][e]
    return (s + RHO[R] - SIG[C]) % 16


def incoming(C):
    return [(R, e) for R in range(4) for e in range(8) if EDGE[R][e][0] == C]


def test_table_expands_to_the_ccsds_matrix(code):
    H = np.zeros((64, 128), np.uint8)
    for R in range(4):
        for e in range(8):
            C, s = EDGE[R][e]
            for i in range(16):
                H[16 * R + i, 16 * C + (i + s) % 16] += 1
    assert np.array_equal(H, np.asarray(code.H).astype(np.uint8))


def test_check_side_gathers_follow_H(code):
    H = np.asarray(code.H)
    for lane in range(16):
        for R in range(4):
            c = 16 * R + (lane + RHO[R]) % 16
            seen = set()
            for e in range(8):
                C = EDGE[R][e][0]
                src = (lane + delta(R, e)) % 16          # rot16(T[C], li + d)
                v = 16 * C + (src + SIG[C]) % 16         # the variable of block C that lane `src` owns
                assert H[c, v] == 1
                seen.add(v)
            assert seen == set(np.flatnonzero(H[c]))


def test_variable_side_sums_in_ascending_check_order(code):
    H = np.asarray(code.H)
    zero_classes = sum(delta(R, e) == 0 for R in range(4) for e in range(8))
    assert zero_classes == 14
    for lane in range(16):
        for C in range(8):
            t = (lane + SIG[C]) % 16
            v = 16 * C + t
            inc = incoming(C)
            assert len(inc) == (5 if C < 4 else 3)
            checks = []
            for (R, e) in inc:
                src = (lane + 16 - delta(R, e)) % 16      # rot16(cv[R][e], li + 16 - d)
                c = 16 * R + (src + RHO[R]) % 16          # the check of block row R that lane `src` owns
                assert H[c, v] == 1
                assert 16 * EDGE[R][e][0] + ((c % 16) + EDGE[R][e][1]) % 16 == v   # and it is THIS edge of that check
                checks.append(c)
            assert sorted(checks) == sorted(np.flatnonzero(H[:, v]))
            # the kernel's ordering fix: the two same-row edges are swapped unless first_lo
            k = next(i for i in range(len(inc) - 1) if inc[i][0] == inc[i + 1][0]) if C < 4 else None
            if k is not None:
                s1, s2 = EDGE[inc[k][0]][inc[k][1]][1], EDGE[inc[k + 1][0]][inc[k + 1][1]][1]
                first_lo = ((t - s1) % 16) < ((t - s2) % 16)
                if C >= 1 and not first_lo:
                    checks[k], checks[k + 1] = checks[k + 1], checks[k]
                if C == 0:
                    assert k == 0  # first two terms of the sum: fp32 addition commutes, no swap needed
                    checks[:2] = sorted(checks[:2])
            assert checks == sorted(checks), (lane, C, checks)


def test_hard_bit_fields_rotate_back_to_code_positions():
    for C in range(8):
        for lane in range(16):
            b = 1 << lane                                   # ballot bit of lane `lane`
            s = SIG[C]
            fld = ((b << s) | (b >> ((16 - s) & 15))) & 0xFFFF
            assert fld == 1 << ((lane + s) % 16)            # = offset of the variable inside block C
