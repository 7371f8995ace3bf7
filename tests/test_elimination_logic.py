"""CPU restatement of the elimination the OSD kernels run (osd_prepare.cuh step 2), checked against the reference's
own elimination (oracle.osd_oracle.swapped_info = swapped_info / identify_mrb / full_gf2elim, PB_OSD/pb_testing.py:231-320).

The kernel does not scan the sorted columns one by one: unit columns of G among the 64 most reliable positions join
the basis on their own row without a visit, the other columns of those positions are visited in order and prefer pivot
rows that no such unit column owns, and positions 64.. are scanned as usual.  This file restates exactly that control
flow on Python integers (one 64-bit word per column, as on the device) and asserts that the basis, the pivot rows'
reduced matrix [I | P'] and the permutation equal the reference's for AWGN frames, frames with ties, and generators
with 64, 32 and 0 unit columns."""
import numpy as np
import pytest

from oracle import osd_oracle as OO
from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

N, K = 128, 64


def kernel_elimination(cols, unit_flag):
    """cols[p]: column at sorted position p (bit r = row r); unit_flag[p]: the handle's "treat as unit column" flag.
    Returns (pivot_row_of_position dict, final columns).  Mirrors osd_prepare.cuh: words 0/1 = positions 0..63."""
    cols = list(cols)
    info = list(unit_flag)
    own = 0
    for p in range(64):
        if info[p]:
            own |= cols[p]
    used = 0            # pivot rows of the visited columns
    av = ~own & (2**64 - 1)
    pivot = {}
    visits = rowops = exact_steps = 0

    def row_op(bit, col_val):
        nonlocal rowops
        m = col_val ^ bit
        rowops += 1
        for q in range(N):
            if cols[q] & bit:
                cols[q] ^= m

    todo = sorted(p for p in range(64) if not info[p])
    while todo:
        p = todo.pop(0)
        visits += 1
        c = cols[p]
        a = c & av
        exact = False
        if a == 0:
            before = 0
            for q in range(p):
                if info[q]:
                    before |= cols[q]
            a = c & ~(used | before) & (2**64 - 1)
            if a == 0:
                continue  # dependent on more reliable columns
            exact = True
            exact_steps += 1
        bit = a & -a
        used |= bit
        av &= ~bit
        row_op(bit, c)
        pivot[p] = bit
        if exact:
            for q in range(p + 1, 64):
                if info[q] and cols[q] & bit:  # a later unit column lost its row: an ordinary column from now on
                    info[q] = False
                    todo.append(q)
            todo.sort()
    for p in range(64):
        if info[p]:
            pivot[p] = cols[p]
    used |= own
    npiv = bin(used).count("1")
    for p in range(64, N):
        if npiv >= K:
            break
        c = cols[p]
        a = c & ~used
        if a == 0:
            continue
        bit = a & -a
        used |= bit
        if c ^ bit:
            row_op(bit, c)
        pivot[p] = bit
        npiv += 1
    return pivot, cols, dict(visits=visits, rowops=rowops, exact=exact_steps)


def run_frame(y, G, unit_of_col):
    pi1 = OO.reliability_order(y)
    gcol = [int(sum(int(G[r, j]) << r for r in range(K))) for j in range(N)]
    pivot, cols, stats = kernel_elimination([gcol[j] for j in pi1], [unit_of_col[j] for j in pi1])
    mrb = sorted(pivot)
    lrb = [p for p in range(N) if p not in pivot]
    assert len(mrb) == K
    # P' row of MRB position t = row `pivot row` of the final matrix restricted to the LRB columns
    red = np.zeros((K, N), dtype=np.int64)
    for t, p in enumerate(mrb):
        r = pivot[p].bit_length() - 1
        red[t, t] = 1
        for u, q in enumerate(lrb):
            red[t, K + u] = (cols[q] >> r) & 1
    perm = pi1[np.array(mrb + lrb)]
    return perm, red, stats


def unit_flags(G):
    seen, out = 0, []
    for j in range(N):
        c = int(sum(int(G[r, j]) << r for r in range(K)))
        unit = c != 0 and c & (c - 1) == 0 and not (seen & c)
        if unit:
            seen |= c
        out.append(unit)
    return out


@pytest.mark.parametrize("mixed_rows", [0, 32, 64])
def test_kernel_elimination_equals_the_reference(mixed_rows):
    code = Code()
    G = np.asarray(code.G, dtype=np.uint8)
    rng = np.random.default_rng(7 + mixed_rows)
    if mixed_rows:
        while True:
            T = np.eye(K, dtype=np.uint8)
            T[:mixed_rows, :mixed_rows] = rng.integers(0, 2, (mixed_rows, mixed_rows), dtype=np.uint8)
            G2 = (T.astype(np.int64) @ G.astype(np.int64) % 2).astype(np.uint8)
            if len(OO.greedy_mrb(rng.normal(size=N).astype(np.float32), G2)[1]) == K:
                G = G2
                break
    flags = unit_flags(G)
    assert sum(flags) == K - mixed_rows
    sigma = float(np.sqrt(1.0 / (2 * 0.5 * 10 ** 0.25)))
    tot = dict(visits=0, rowops=0, exact=0)
    n = 60
    for i in range(n):
        y = (1.0 + sigma * rng.standard_normal(N)).astype(np.float32)
        if i % 10 == 9:
            y = np.round(y * 2) / 2  # ties: the stable order decides, many equal reliabilities
        _, _, red_ref, perm_ref = OO.swapped_info(y, np.zeros(N, np.int64), G)
        perm, red, stats = run_frame(y, G, flags)
        assert np.array_equal(perm, perm_ref)
        assert np.array_equal(red, np.asarray(red_ref))
        for k in tot:
            tot[k] += stats[k]
    if mixed_rows == 0:  # the point of the exercise: about half the serial steps of a column-by-column scan (~66)
        assert tot["visits"] / n < 36 and tot["rowops"] / n < 40
