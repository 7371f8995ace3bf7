"""-m gpu parity tests: the CUDA path through the C ABI versus the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): OSD sort / elimination / permutation / TEP choice / decision
bit-exact; NMS hard decisions identical and every soft output within 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import nms_oracle as NO
from oracle import osd_oracle as OO
from oracle import philox_oracle as PO
from short_ldpc_decoding_osd_b200 import _lib
from tests.gpu_util import dev, empty, nms_gpu, osd_gpu, redG_to_matrix, sync

pytestmark = pytest.mark.gpu
ALPHA = float(NO.softplus(-0.048))


def frames(code, B, ebn0=2.5, seed=0, first=0):
    return PO.gen_frames(seed, first, B, ebn0, code.G)


# ---- frame generator -----------------------------------------------------------------------------
@pytest.mark.parametrize("seed,first,B,ebn0", [(0, 0, 257, 2.5), (12345678901234, 2**33 + 5, 64, 4.0), (7, 0, 1, 1.5)])
def test_framegen_matches_oracle(handle, code, seed, first, B, ebn0):
    y, cw, _ = PO.gen_frames(seed, first, B, ebn0, code.G)
    yd = empty((B, 128), torch.float32)
    cd = empty((B, 4), torch.int32)
    handle.call("ldpcb_gen_frames", seed, first, B, float(ebn0), yd, cd, None)
    sync()
    got_cw = _lib.unpack_bits(cd.cpu().numpy().view(np.uint32))
    assert np.array_equal(got_cw, cw)  # message bits and codeword: bit-exact
    # fp32 logf/sqrtf/sincospif vs float64 oracle: a few ulp of |y| <= ~6
    np.testing.assert_allclose(yd.cpu().numpy(), y, rtol=0, atol=4e-6)


def test_framegen_sharding_is_position_independent(handle, code):
    B = 300
    full = empty((B, 128), torch.float32)
    part = empty((100, 128), torch.float32)
    handle.call("ldpcb_gen_frames", 3, 1000, B, 2.5, full, None, None)
    handle.call("ldpcb_gen_frames", 3, 1100, 100, 2.5, part, None, None)
    sync()
    assert torch.equal(full[100:200], part)


def test_framegen_statistics(handle, code):
    B = 200000
    yd = empty((B, 128), torch.float32)
    cd = empty((B, 4), torch.int32)
    handle.call("ldpcb_gen_frames", 99, 0, B, 2.5, yd, cd, None)
    sync()
    cw = torch.from_numpy(_lib.unpack_bits(cd.cpu().numpy().view(np.uint32)).astype(np.float32)).to(yd.device)
    z = (yd * (1 - 2 * cw) - 1.0) / float(PO.sigma_of(2.5))
    zz = z.double()
    n = zz.numel()
    assert abs(zz.mean().item()) < 5 / np.sqrt(n)
    assert abs(zz.var().item() - 1) < 5 * np.sqrt(2 / n)
    assert abs((zz**4).mean().item() - 3) < 5 * np.sqrt(96 / n)
    assert abs(cw.mean().item() - 0.5) < 5 * 0.5 / np.sqrt(n)


# ---- NMS -------------------------------------------------------------------------------------------
def check_nms(got, ref, iters):
    assert np.array_equal(got["hard"], ref["hard"])
    assert np.array_equal(got["syndrome_nz"], ref["syndrome_nz"])
    assert np.array_equal(got["iters_used"], ref["iters_used"])
    # float messages within 1e-5 relative (north_star); atol covers posteriors that cancel to ~0
    np.testing.assert_allclose(got["traj"], ref["traj"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("ebn0,B", [(2.5, 700), (1.0, 200), (5.0, 200)])
def test_nms_matches_oracle(handle, code, ebn0, B):
    y, cw, _ = frames(code, B, ebn0, seed=11)
    ref = NO.decode(y, code.H, 12, ALPHA)
    got = nms_gpu(handle, y, 12, ALPHA)
    check_nms(got, ref, 12)


def test_nms_bit_weights_and_iters(handle, code):
    y, _, _ = frames(code, 150, 2.0, seed=5)
    for iters, a, wv, wm in [(5, 0.8, 0.9, 1.1), (1, 0.5, 1.0, 1.0), (20, 0.7, 1.05, 1.05)]:
        ref = NO.decode(y, code.H, iters, a, wv, wm)
        got = nms_gpu(handle, y, iters, a, wv, wm)
        check_nms(got, ref, iters)


def test_nms_early_stop(handle, code):
    y, _, _ = frames(code, 400, 2.5, seed=21)
    ref = NO.decode(y, code.H, 12, ALPHA, early_stop=True)
    got = nms_gpu(handle, y, 12, ALPHA, early=1)
    check_nms(got, ref, 12)
    assert got["iters_used"].min() < 12


def test_nms_edge_inputs(handle, code):
    _, cw, _ = frames(code, 6, 2.5, seed=2)
    y = (1.0 - 2.0 * cw).astype(np.float32)  # noiseless
    y[1] = 0.0                                # all-zero LLRs: tf.sign(0)=0 path, hard decision = all ones
    y[2, :5] = 0.0                            # a few exact zeros
    y[3] *= 1e-30
    y[4] *= 1e20
    y[5, 7] = -0.0
    ref = NO.decode(y, code.H, 12, ALPHA)
    got = nms_gpu(handle, y, 12, ALPHA)
    assert np.array_equal(got["hard"], ref["hard"])
    assert np.array_equal(got["syndrome_nz"], ref["syndrome_nz"])
    np.testing.assert_allclose(got["traj"], ref["traj"], rtol=1e-5, atol=1e-37)
    assert np.array_equal(got["hard"][0], cw[0])


def test_nms_empty_and_ragged(handle, code):
    handle.call("ldpcb_nms_decode", None, 0, 12, ALPHA, 1.0, 1.0, 0, None, None, None, None, None)
    for B in (1, 7, 9, 1187):
        y, _, _ = frames(code, B, 2.5, seed=B)
        ref = NO.decode(y, code.H, 12, ALPHA)
        got = nms_gpu(handle, y, 12, ALPHA, traj=False)
        assert np.array_equal(got["hard"], ref["hard"])


def test_nms_rejects_bad_arguments(handle):
    y = empty((4, 128), torch.float32)
    bits = empty((4, 4), torch.int32)
    with pytest.raises(_lib.LdpcB200Error) as e:
        handle.call("ldpcb_nms_decode", y.data_ptr() + 4, 3, 12, ALPHA, 1.0, 1.0, 0, bits, None, None, None, None)
    assert e.value.status == -3
    with pytest.raises(_lib.LdpcB200Error) as e:
        handle.call("ldpcb_nms_decode", y, 4, 999, ALPHA, 1.0, 1.0, 0, bits, None, None, None, None)
    assert e.value.status == -1


# ---- OSD -------------------------------------------------------------------------------------------
def failing_frames(code, n, ebn0=2.5, seed=31):
    y, cw, _ = frames(code, max(6 * n, 200), ebn0, seed=seed)
    ref = NO.decode(y, code.H, 12, ALPHA)
    idx = np.flatnonzero(ref["syndrome_nz"])[:n]
    return y[idx], cw[idx]


def check_osd_frame(got, i, ref, code):
    assert np.array_equal(got["perm"][i], ref["perm"])
    assert np.array_equal(redG_to_matrix(got["redG"][i]), ref["reduced_G"])
    assert int(got["score_exp"][i]) == ref["score_exp"]
    assert int(got["best_tep"][i]) == ref["best_tep"]
    assert int(got["best_score_q"][i]) == ref["best_score_q"]
    assert np.array_equal(got["codeword"][i], ref["codeword"])


@pytest.mark.parametrize("order,n", [(0, 40), (1, 40), (2, 24)])
def test_osd_conv_bit_exact(handle, code, order, n):
    y, cw = failing_frames(code, n)
    teps = OO.generate_teps_conv(order)
    got = osd_gpu(handle, y, order=order)
    for i in range(len(y)):
        check_osd_frame(got, i, OO.osd_frame(y[i], y[i], code.G, teps), code)
    # every output is a codeword
    assert not (got["codeword"].astype(np.int64).dot(code.H.T) % 2).any()


def test_osd_order3_bit_exact(handle, code):
    y, cw = failing_frames(code, 3, seed=77)
    teps = OO.generate_teps_conv(3)
    got = osd_gpu(handle, y, order=3)
    for i in range(len(y)):
        check_osd_frame(got, i, OO.osd_frame(y[i], y[i], code.G, teps), code)


def test_osd_fs_order_and_flags(handle, code):
    y, cw = failing_frames(code, 16, seed=41)
    rng = np.random.default_rng(5)
    ys = (y + rng.normal(0, 0.3, y.shape)).astype(np.float32)  # a different scoring metric (DL path shape)
    teps = OO.generate_teps_fs(2)
    for flags in (0, 1, 2, 3):
        got = osd_gpu(handle, y, ys, order=2, tep_order=1, flags=flags)
        for i in range(len(y)):
            check_osd_frame(got, i, OO.osd_frame(y[i], ys[i], code.G, teps, flags=flags), code)


def test_osd_ties_and_degenerate_inputs(handle, code):
    y, cw = failing_frames(code, 8, seed=51)
    y = y.copy()
    y[0] = np.round(y[0] * 4) / 4          # many equal |y|: exercises the stable tie rule
    y[1] = np.where(cw[1] == 0, 1.0, -1.0)  # all |y| equal
    y[2] = 0.0                              # all zero: hard decision all ones, scores all zero
    y[3, ::3] = 0.0
    y[4] *= 1e-38                           # denormals
    y[5] *= 1e30
    teps = OO.generate_teps_conv(1)
    for flags in (0, 1):
        got = osd_gpu(handle, y, order=1, flags=flags)
        for i in range(len(y)):
            check_osd_frame(got, i, OO.osd_frame(y[i], y[i], code.G, teps, flags=flags), code)


def test_osd_fixes_what_order_p_can_fix(handle, code):
    """'Miracle view' bound (PB_OSD/pb_testing.py:502-511): if the MRB hard decision has <= p errors
    the order-p result scores no worse than the transmitted codeword."""
    y, cw = failing_frames(code, 64, seed=61)
    got = osd_gpu(handle, y, order=2)
    for i in range(len(y)):
        perm = got["perm"][i]
        hard = (~(y[i][perm] > 0)).astype(np.int64)
        mrb_err = int((hard[:64] ^ cw[i][perm][:64]).sum())
        q, _ = OO.quantize(y[i][perm])
        truth_score = int(q[(cw[i][perm] ^ hard) == 1].sum())
        if mrb_err <= 2:
            assert int(got["best_score_q"][i]) <= truth_score


def test_osd_block_minima(handle, code):
    y, cw = failing_frames(code, 10, seed=71)
    rng = np.random.default_rng(9)
    yo = (y + rng.normal(0, 0.2, y.shape)).astype(np.float32)  # ordering metric != channel (DL shape)
    sizes, bnd = OO.dl_segments()
    ranges = [range(bnd[i], bnd[i + 1]) for i in range(6)]
    path = [[0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 2, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0], [0, 1, 1, 0, 1, 0]]
    blocks = [OO.dl_error_pattern_block(p, ranges) for p in path]
    teps_dl = [t for b in blocks for t in b]
    starts = np.concatenate([[0], np.cumsum([len(b) for b in blocks])]).astype(np.int32)
    packed = OO.pack_teps(teps_dl, dl_index=True)
    teps_mr = [tuple(sorted(63 - x for x in t)) for t in teps_dl]
    B = len(y)
    flags = OO.TIES_HIGH_INDEX_FIRST | OO.DISC_HARD_FROM_SCORE
    bm = empty((B, len(blocks)), torch.int64)
    ba = empty((B, len(blocks)), torch.int32)
    ex = empty((B,), torch.int32)
    ts = empty((B,), torch.int64)
    pm = empty((B, 128), torch.uint8)
    handle.call("ldpcb_osd_block_minima", dev(yo), dev(y), B, dev(packed.view(np.int32)), len(packed), dev(starts), len(blocks), flags,
                bm, ba, ex, dev(_lib.pack_bits(cw).view(np.int32)), ts, pm, None)
    sync()
    for i in range(B):
        ref = OO.osd_frame(yo[i], y[i], code.G, teps_mr, flags=flags, block_start=starts, truth=cw[i])
        assert np.array_equal(pm[i].cpu().numpy(), ref["perm"])
        assert np.array_equal(bm[i].cpu().numpy(), ref["block_min_q"])
        assert np.array_equal(ba[i].cpu().numpy(), ref["block_arg"])
        assert int(ts[i]) == ref["truth_score_q"]
        assert int(ex[i]) == ref["score_exp"]


def test_tep_tables_match_oracle(handle):
    for order, n in [(0, 1), (1, 65), (2, 2081), (3, 43745)]:
        assert handle.tep_count(order, 0) == n and handle.tep_count(order, 1) == n
        assert np.array_equal(handle.tep_table(order, 0), OO.pack_teps(OO.generate_teps_conv(order)))
        assert np.array_equal(handle.tep_table(order, 1), OO.pack_teps(OO.generate_teps_fs(order)))


# ---- another (128,64) code: the generic instantiations ---------------------------------------------------------
def _other_code():
    """H = [A | I], G = [I | A^T] with A a sum of six or seven 64x64 circulant permutations: check degrees 7 and 8
    (padded edge slots), variable degrees 6..7 and 1 -- none of the CCSDS fast-path assumptions hold."""
    A = np.zeros((64, 64), dtype=np.uint8)
    for s in (1, 5, 11, 24, 37, 50):
        A[np.arange(64), (np.arange(64) + s) % 64] = 1
    rows = np.arange(0, 64, 3)
    A[rows, (rows + 58) % 64] = 1
    H = np.concatenate([A, np.eye(64, dtype=np.uint8)], axis=1)
    G = np.concatenate([np.eye(64, dtype=np.uint8), A.T], axis=1)
    assert not ((H.astype(np.int64) @ G.T.astype(np.int64)) % 2).any()
    assert H.sum(1).max() <= 8 and H.sum(0).max() <= 8
    return H, G


def test_other_code_nms_and_osd(code):
    H, G = _other_code()
    h = _lib.Handle(H, G, device=0)
    try:
        y, cw, _ = PO.gen_frames(21, 0, 600, 3.0, G)
        assert not ((cw.astype(np.int64) @ H.T.astype(np.int64)) % 2).any()
        ref = NO.decode(y, H, 12, ALPHA)
        got = nms_gpu(h, y, 12, ALPHA)
        check_nms(got, ref, 12)
        for w_vc, w_marg in ((0.8, 1.1),):
            ref2 = NO.decode(y[:100], H, 7, 0.75, w_vc, w_marg)
            got2 = nms_gpu(h, y[:100], 7, 0.75, w_vc, w_marg)
            check_nms(got2, ref2, 7)
        fails = np.flatnonzero(ref["syndrome_nz"])[:12]
        assert len(fails) >= 4
        for order in (1, 2):
            teps = OO.generate_teps_conv(order)
            o = osd_gpu(h, y[fails], order=order)
            for i, f in enumerate(fails):
                check_osd_frame(o, i, OO.osd_frame(y[f], y[f], G, teps), code)
        # the whole pipeline on this code: decisions of detected failures replaced by order-2 OSD
        B = len(y)
        bits = np.empty((B, 4), np.uint32)
        syn = np.empty(B, np.uint8)
        cnt = np.zeros(16, np.uint64)
        h.call("ldpcb_decode_host", y, B, 12, ALPHA, 1.0, 1.0, 0, 2, 0, bits, syn, None, _lib.pack_bits(cw), cnt)
        want = ref["hard"].copy()
        allf = np.flatnonzero(ref["syndrome_nz"])
        want[allf] = CO.osd(np.ascontiguousarray(y[allf]), None, G, OO.pack_teps(OO.generate_teps_conv(2)))["codeword"]
        assert np.array_equal(_lib.unpack_bits(bits), want)
        assert int(cnt[0]) == B and int(cnt[9]) == int((want != cw).any(1).sum())
    finally:
        h.close()


@pytest.mark.parametrize("mixed_rows", [64, 32, 5])
def test_osd_does_not_depend_on_the_generator_form(handle, code, mixed_rows):
    """The OSD elimination treats the unit columns of G specially (osd_prepare.cuh step 2).  The MRB, the reduced
    matrix and the decision depend on the code and the frame only, so a handle built on T.G -- T invertible, mixing
    `mixed_rows` generator rows, which leaves 64 - mixed_rows unit columns (none for 64) -- must reproduce every output
    of the systematic handle bit for bit; a few frames are also checked against the oracle run on T.G."""
    rng = np.random.default_rng(100 + mixed_rows)
    G = np.asarray(code.G, dtype=np.uint8)
    while True:
        T = np.eye(64, dtype=np.uint8)
        T[:mixed_rows, :mixed_rows] = rng.integers(0, 2, (mixed_rows, mixed_rows), dtype=np.uint8)
        G2 = (T.astype(np.int64) @ G.astype(np.int64) % 2).astype(np.uint8)
        if _gf2_rank(G2) == 64:
            break
    unit_cols = int(((G2.sum(0) == 1)).sum())
    assert unit_cols <= 64 - mixed_rows + 2
    y, cw = failing_frames(code, 1500, seed=91)
    y = np.ascontiguousarray(y)
    h2 = _lib.Handle(code.H, G2, device=0)
    try:
        for order, flags in ((1, 0), (2, 0), (2, 1)):
            a = osd_gpu(handle, y, order=order, flags=flags)
            b = osd_gpu(h2, y, order=order, flags=flags)
            for key in ("perm", "redG", "score_exp", "best_tep", "best_score_q", "codeword"):
                assert np.array_equal(a[key], b[key]), (key, order, flags)
        teps = OO.generate_teps_conv(1)
        b = osd_gpu(h2, y[:6], order=1)
        for i in range(6):
            check_osd_frame(b, i, OO.osd_frame(y[i], y[i], G2, teps), code)
    finally:
        h2.close()


def _gf2_rank(M):
    A = np.array(M, dtype=np.uint8) & 1
    r = 0
    for c in range(A.shape[1]):
        p = np.flatnonzero(A[r:, c])
        if len(p) == 0:
            continue
        p = p[0] + r
        A[[r, p]] = A[[p, r]]
        hit = np.flatnonzero(A[:, c])
        hit = hit[hit != r]
        A[hit] ^= A[r]
        r += 1
        if r == A.shape[0]:
            break
    return r
