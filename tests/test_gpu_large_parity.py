"""-m gpu: parity at scale against the C oracle (bit-identical to the NumPy oracle, see test_oracle_c.py),
and FER against the reference points of tests/golden/fer_reference.json."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import osd_oracle as OO
from oracle import philox_oracle as PO
from short_ldpc_decoding_osd_b200 import _lib, simulate
from tests.gpu_util import dev, empty, nms_gpu, osd_gpu, sync

pytestmark = pytest.mark.gpu
ALPHA = 0.66943514


def test_nms_200k_frames_bit_exact(handle, code):
    B = 200_000
    y, cw, _ = PO.gen_frames(31337, 0, B, 2.5, code.G)
    ref = CO.nms(y, code.H, 12, ALPHA)
    got = nms_gpu(handle, y, 12, ALPHA, traj=False)
    agree = (got["hard"] == ref["hard"]).all(axis=1).mean()
    assert agree == 1.0  # bar: >= 99.99% of frames
    assert np.array_equal(got["syndrome_nz"], ref["syndrome_nz"])
    sub = slice(0, 20000)
    ref_t = CO.nms(y[sub], code.H, 12, ALPHA, traj=True)["traj"]
    got_t = nms_gpu(handle, y[sub], 12, ALPHA)["traj"]
    np.testing.assert_allclose(got_t, ref_t, rtol=1e-5, atol=2e-6)
    assert (got_t == ref_t).mean() > 0.9999  # in practice bit-identical
    # early-stop variant too
    ref_e = CO.nms(y[sub], code.H, 12, ALPHA, early_stop=True)
    got_e = nms_gpu(handle, y[sub], 12, ALPHA, early=1, traj=False)
    assert np.array_equal(got_e["hard"], ref_e["hard"]) and np.array_equal(got_e["iters_used"], ref_e["iters_used"])


@pytest.mark.parametrize("order,n,tep_order,flags", [(1, 40000, 0, 0), (2, 20000, 0, 0), (2, 6000, 1, 1), (3, 600, 0, 0), (3, 3000, 1, 1), (3, 1501, 0, 0)])
def test_osd_many_frames_bit_exact(handle, code, order, n, tep_order, flags):
    y, cw, _ = PO.gen_frames(4242 + order, 0, 6 * n, 2.5, code.G)
    syn = CO.nms(y, code.H, 12, ALPHA)["syndrome_nz"]
    yf = np.ascontiguousarray(y[syn][:n])
    teps = OO.pack_teps(OO.generate_teps_conv(order) if tep_order == 0 else OO.generate_teps_fs(order))
    ref = CO.osd(yf, None, code.G, teps, flags=flags)
    got = osd_gpu(handle, yf, order=order, tep_order=tep_order, flags=flags)
    for k in ("perm", "best_tep", "best_score_q", "score_exp", "codeword", "redG"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("order,n", [(1, 8000), (2, 3000), (3, 400)])
def test_osd_quantised_inputs_many_ties(handle, code, order, n):
    """Every frame has many equal |y|: the exact tie path of the sort and, at order 2, the exact 64-bit fallback
    of the tensor-core pair sweep (many candidates share the truncated minimum), at scale."""
    y, cw, _ = PO.gen_frames(99, 0, 30000, 2.0, code.G)
    yq = (np.round(y * 8) / 8).astype(np.float32)
    teps = OO.pack_teps(OO.generate_teps_conv(order))
    for flags in (0, 1):
        ref = CO.osd(yq[:n], None, code.G, teps, flags=flags)
        got = osd_gpu(handle, yq[:n], order=order, flags=flags)
        for k in ("perm", "best_tep", "best_score_q", "codeword", "redG"):
            assert np.array_equal(got[k], ref[k]), (flags, k)


def test_fer_matches_reference_points(handle, golden_dir):
    """FER of the GPU pipeline (Philox frames) against the reference generator + reference decoder points.
    Two independent Monte-Carlo estimates: |difference| must stay below 3.5 combined standard errors, and the
    inside-the-95%-CI flags are reported."""
    ref = json.load(open(os.path.join(golden_dir, "fer_reference.json")))
    inside = []
    for i, pt in enumerate(ref["points"]):
        for order in (1, 2):
            n = 1 << 22
            t = simulate.run_point(handle, pt["ebn0_db"], n, seed=500 + i, osd_order=order, chunk=1 << 21)
            for ours, theirs, ci, k_ref in ((t.fer_nms, pt["fer_nms"], pt["fer_nms_ci95"], pt["frames"]),
                                            (t.fer_final, pt[f"fer_final_osd{order}"], pt[f"fer_final_osd{order}_ci95"], pt["frames"])):
                se = np.sqrt(ours * (1 - ours) / n + theirs * (1 - theirs) / k_ref)
                assert abs(ours - theirs) <= 3.5 * se + 1e-12, (pt["ebn0_db"], order, ours, theirs, se)
                inside.append(ci[0] <= ours <= ci[1])
    assert np.mean(inside) >= 0.8  # a 95% interval of an independent estimate is missed ~5% of the time by chance


@pytest.mark.parametrize("tep_order,ebn0,flags,other_metric", [(0, 2.0, 0, False), (1, 3.0, 0, False), (0, 2.5, 3, True), (1, 2.5, 1, True)])
def test_pair_sweep_equals_lut_sweep_on_a_million_frames(handle, tep_order, ebn0, flags, other_metric):
    """The tensor-core pair sweep (order-2 lists, truncated scores + exact re-scoring of the window) against the
    exact 64-bit byte-LUT sweep of the same kernel family (block-minima path with LDPCB_BLOCKS_LUT=1, one block = the whole list) on 2^20
    device-generated frames: first-minimum index and exact score must agree on every frame.  With `other_metric` the
    scoring metric differs from the ordering metric (the DL path's shape: negative per-position deltas) and the tie /
    discrepancy flags are set."""
    B = 1 << 20
    y = empty((B, 128), torch.float32)
    tr = empty((B, 4), torch.int32)
    handle.call("ldpcb_gen_frames", 9091 + tep_order, 0, B, float(ebn0), y, tr, None)
    ys = y
    if other_metric:
        g = torch.Generator(device=y.device)
        g.manual_seed(5)
        ys = (y + 0.3 * torch.randn(y.shape, generator=g, device=y.device, dtype=torch.float32)).contiguous()
    cw = empty((B, 4), torch.int32)
    bt = empty((B,), torch.int32)
    bq = empty((B,), torch.int64)
    ex = empty((B,), torch.int32)
    handle.call("ldpcb_osd_decode", y, ys, B, 2, tep_order, flags, cw, bt, bq, ex, None, None, None)
    teps = handle.tep_table(2, tep_order)
    starts = np.array([0, len(teps)], np.int32)
    bm = empty((B, 1), torch.int64)
    ba = empty((B, 1), torch.int32)
    ex2 = empty((B,), torch.int32)
    import os

    os.environ["LDPCB_BLOCKS_LUT"] = "1"
    try:
        handle.call("ldpcb_osd_block_minima", y, ys, B, dev(teps.view(np.int32)), len(teps), dev(starts), 1, flags, bm, ba, ex2, None, None, None, None)
        sync()
    finally:
        os.environ.pop("LDPCB_BLOCKS_LUT", None)
    assert torch.equal(ex, ex2)
    assert torch.equal(bq, bm[:, 0])
    assert torch.equal(bt, ba[:, 0])


@pytest.mark.parametrize("threshold_sum,quantised", [(2, False), (2, True), (3, False), (3, True)])
def test_block_minima_truncated_sweep_equals_lut_sweep_on_many_frames(code, threshold_sum, quantised):
    """osd_blocks.cu (per-block truncated shuffle-table sweep + exact re-scoring of the window, marked blocks redone
    through the byte LUT) against the byte-LUT sweep of round 1 (LDPCB_BLOCKS_LUT=1) on the NMS failures of 2^19 frames:
    every block minimum and its first index, the transmitted codeword's score, exponent and permutation.  Blocks: the
    DL order patterns over the six segments with total weight <= threshold_sum (27 / 77 blocks, 2081 / 43745 TEPs);
    quantised LLRs make near-ties the rule, so most blocks go through the marked-block redo."""
    import os

    import torch

    from short_ldpc_decoding_osd_b200 import _lib, globalmap as GL, nn_testing
    from short_ldpc_decoding_osd_b200 import ordered_statistics_decoding as OSD
    from tests.gpu_util import dev, empty, sync

    for k, v in dict(code_parameters=code, threshold_sum=threshold_sum, segment_num=6, decoding_length=100).items():
        GL.set_map(k, v)
    teps_list, acc = nn_testing.generate_teps(OSD.osd(code), nn_testing.filter_order_patterns(nn_testing.convention_segment_path()))
    packed = np.concatenate([OSD.pack_dl_teps(b) for b in teps_list])
    nb = len(teps_list)
    assert (nb, len(packed)) == ((27, 2081) if threshold_sum == 2 else (77, 43745))
    h = _lib.Handle(code.H, code.G, device=0)
    B = 1 << 19 if threshold_sum == 2 else 1 << 17
    y = empty((B, 128), torch.float32)
    tr = empty((B, 4), torch.int32)
    h.call("ldpcb_gen_frames", 77, 0, B, 2.5, y, tr, None)
    if quantised:
        y = (torch.round(y * 2.0) / 2.0).contiguous()
    bits, syn, met = empty((B, 4), torch.int32), empty((B,), torch.uint8), empty((B, 128), torch.float32)
    h.call("ldpcb_nms_decode_fir", y, B, 12, 0.66943514, 1.0, 1.0, np.full(13, 1 / 13, np.float32), 0.0, bits, syn, met, None)
    m = syn.bool()
    yf, mf, tf = y[m].contiguous(), met[m].contiguous(), tr[m].contiguous()
    n = int(yf.shape[0])
    assert n > 1000
    out = {}
    for tag in ("sweep32", "lut"):
        if tag == "lut":
            os.environ["LDPCB_BLOCKS_LUT"] = "1"
        try:
            bm, ba, ex, ts, pm = empty((n, nb), torch.int64), empty((n, nb), torch.int32), empty((n,), torch.int32), empty((n,), torch.int64), empty((n, 128), torch.uint8)
            h.call("ldpcb_osd_block_minima", mf, yf, n, dev(packed.view(np.int32)), len(packed), dev(np.asarray(acc, np.int32)), nb,
                   OSD.FLAGS_DL | (threshold_sum << _lib.OSD_MAXW_SHIFT), bm, ba, ex, tf, ts, pm, None)
            sync()
            out[tag] = [x.cpu().numpy() for x in (bm, ba, ex, ts, pm)]
        finally:
            os.environ.pop("LDPCB_BLOCKS_LUT", None)
    for name, u, v in zip(("block_min_q", "block_arg", "score_exp", "truth_score_q", "perm"), out["sweep32"], out["lut"]):
        assert np.array_equal(u, v), (name, int((u != v).sum()), n)
    assert (out["sweep32"][0] >= 0).all()
    h.close()


def test_order3_tensor_sweep_equals_generic_sweep_on_many_frames(code):
    """osd3.cu (62 pair problems on IMMA + exact re-scoring) against the generic order-3 kernel (32-bit shuffle-table
    sweep) on 2^17 frames of a low-SNR batch: decisions, TEP choices and exact scores, both TEP orders, DL flags."""
    import os

    import torch

    from short_ldpc_decoding_osd_b200 import _lib
    from tests.gpu_util import dev, empty, sync

    hq = _lib.Handle(code.H, code.G, device=0)
    n = 1 << 17
    y = empty((n, 128), torch.float32)
    y2 = empty((n, 128), torch.float32)
    hq.call("ldpcb_gen_frames", 5, 0, n, 1.5, y, None, None)
    hq.call("ldpcb_gen_frames", 6, 0, n, 1.5, y2, None, None)
    out = {}
    for tag in ("tensor", "generic"):
        if tag == "generic":
            os.environ["LDPCB_OSD3_GENERIC"] = "1"
        try:
            for tep_order, flags, ys in ((0, 0, y), (1, 3, y2)):
                cw, bt, bq = empty((n, 4), torch.int32), empty((n,), torch.int32), empty((n,), torch.int64)
                hq.call("ldpcb_osd_decode", y, ys, n, 3, tep_order, flags, cw, bt, bq, None, None, None, None)
                sync()
                out[(tag, tep_order)] = (cw.cpu().numpy(), bt.cpu().numpy(), bq.cpu().numpy())
        finally:
            os.environ.pop("LDPCB_OSD3_GENERIC", None)
    for tep_order in (0, 1):
        for u, v, name in zip(out[("tensor", tep_order)], out[("generic", tep_order)], ("codeword", "best_tep", "best_score_q")):
            assert np.array_equal(u, v), (tep_order, name, int((u != v).sum()))
    hq.close()


@pytest.mark.parametrize("ebn0,tau_e,tau_psc,beta", [(2.5, 6.5, 30, 6.4), (1.5, 6.5, 30, 6.4), (2.5, 9.0, 26, 12.8), (3.5, 4.0, 34, 3.2)])
def test_fs_order3_three_launch_path_equals_single_kernel(code, ebn0, tau_e, tau_psc, beta):
    """FS policy at order_limit 3: classes 0..2 + deferred tensor-core class 3 + exact kernel for the undecided frames
    (launch_osd_fs3) against the single exact kernel (LDPCB_FS3_EXACT=1): codeword, decision index, visited TEPs, stop kind
    and exact score of every frame, on NMS failures at several SNRs and threshold settings."""
    import os

    import torch

    from short_ldpc_decoding_osd_b200 import _lib
    from tests.gpu_util import empty, sync

    h = _lib.Handle(code.H, code.G, device=0)
    B = 1 << 18
    y = empty((B, 128), torch.float32)
    h.call("ldpcb_gen_frames", 11, 0, B, ebn0, y, None, None)
    bits, it, syn = empty((B, 4), torch.int32), empty((B,), torch.uint8), empty((B,), torch.uint8)
    h.call("ldpcb_nms_decode", y, B, 12, ALPHA, 1.0, 1.0, 0, bits, it, syn, None, None)
    yf = y[syn.bool()][:40000].contiguous()
    n = yf.shape[0]
    out = {}
    for tag in ("three", "exact"):
        if tag == "exact":
            os.environ["LDPCB_FS3_EXACT"] = "1"
        try:
            cw, bt, nt, sk, bq = empty((n, 4), torch.int32), empty((n,), torch.int32), empty((n,), torch.int32), empty((n,), torch.uint8), empty((n,), torch.int64)
            h.call("ldpcb_osd_fs_decode", yf, n, 3, tau_e, tau_psc, beta, cw, bt, nt, sk, bq, None, None, None)
            sync()
            out[tag] = [x.cpu().numpy() for x in (cw, bt, nt, sk, bq)]
        finally:
            os.environ.pop("LDPCB_FS3_EXACT", None)
    for name, u, v in zip(("codeword", "best_tep", "num_teps", "stop_kind", "best_score_q"), out["three"], out["exact"]):
        assert np.array_equal(u, v), (name, int((u != v).sum()), n)
    if (ebn0, beta) == (2.5, 6.4):
        assert (out["exact"][3] == 3).sum() > 0  # at the reference's settings some frames do sweep all three classes
    h.close()
