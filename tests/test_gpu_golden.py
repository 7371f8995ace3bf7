"""-m gpu: the CUDA path and the drop-in modules against the golden vectors produced by the reference's
own source (tests/golden/*.npz, oracle/ref_runner.py)."""
import os

import numpy as np
import pytest

from oracle import nms_oracle as NO
from oracle import osd_oracle as OO
from short_ldpc_decoding_osd_b200 import _lib
from short_ldpc_decoding_osd_b200 import globalmap as GL
from tests.gpu_util import nms_gpu, osd_gpu, redG_to_matrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def settings(code):
    GL.set_map("code_parameters", code)
    GL.set_map("selected_decoder_type", "NMS-1")
    GL.set_map("num_iterations", 12)
    GL.set_map("order_limit", 2)
    GL.set_map("d_min", 14)
    GL.set_map("tau_psc", 30)
    GL.set_map("termination_num_threshlod", 100)
    for k in ("convention_osd", "miracle_view"):
        GL.set_map(k, False)
    yield


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_nms_kernel_matches_reference_graph(handle, golden_dir):
    g = load(golden_dir, "nms_ref_shim.npz")
    got = nms_gpu(handle, g["y"], 12, float(g["alpha"][0]))
    ref_hard = NO.hard_decision(g["soft"][:, 12]).astype(np.uint8)
    assert np.array_equal(got["hard"], ref_hard)  # 100% of frames identical (bar: >= 99.99%)
    np.testing.assert_allclose(got["traj"], g["soft"], rtol=1e-5, atol=2e-6)  # every message-derived posterior
    assert np.array_equal(np.flatnonzero(got["syndrome_nz"])[:, None], g["index"])


def test_decoding_model_dropin_matches_reference(handle, golden_dir):
    from short_ldpc_decoding_osd_b200 import ms_test

    g = load(golden_dir, "nms_ref_shim.npz")
    model = ms_test.Decoding_model()
    assert np.float32(model.layer.shared_check_weight[0]) == np.float32(g["raw_weight"][0])
    fer, ber, und, buffer = model(g["y"], g["labels"].astype(np.int64))
    assert fer == pytest.approx(float(g["fer"]), abs=1e-12) and ber == pytest.approx(float(g["ber"]), abs=1e-12)
    assert und == int(g["undetected"])
    assert len(buffer[0]) == len(buffer[1]) == int(g["n_buffer_rows"])
    np.testing.assert_allclose(np.stack(buffer[0][:13]), g["buffer_first"], rtol=1e-5, atol=2e-6)
    soft = model.layer(g["y"], g["labels"])
    assert len(soft) == 13 and soft[0].shape == (96, 128)
    fer2, ber2, und2, index = model.get_eval(soft, g["labels"].astype(np.int64))
    assert (fer2, und2) == (fer, und) and np.array_equal(index, g["index"])
    flat = model.postprocess_failure_cases(([buffer[0]], [buffer[1]]))
    assert len(flat[0]) == len(buffer[0])


def test_osd_kernel_matches_reference_swapped_info_and_sweep(handle, golden_dir):
    g = load(golden_dir, "osd_ref_shim.npz")
    for order in (1, 2):
        got = osd_gpu(handle, g["y"], order=order)
        for i in range(len(g["y"])):
            perm = got["perm"][i].astype(np.int64)
            assert np.array_equal(g["y"][i][perm], g["upd_in"][i])            # pi2 o pi1
            assert np.array_equal(g["labels"][i][perm], g["upd_lab"][i])
            assert np.array_equal(redG_to_matrix(got["redG"][i]), np.unpackbits(g["redG"][i], axis=1)[:, :128])
            assert int(got["best_tep"][i]) == int(g[f"idx{order}"][i])         # TEP choice
            assert bool((got["codeword"][i] == g["labels"][i]).all()) == bool(g[f"ok{order}"][i])


def test_osd_dropins_match_reference(handle, golden_dir):
    from short_ldpc_decoding_osd_b200 import convention_osd as C
    from short_ldpc_decoding_osd_b200 import pb_testing as P

    g = load(golden_dir, "osd_ref_shim.npz")
    for o in (0, 1, 2):
        assert np.array_equal(C.generate_teps(o), np.unpackbits(g[f"teps{o}"], axis=1)[:, :64])
    assert C.query_boundary(2) == list(g["boundary2"])
    for i in range(0, len(g["y"]), 3):
        ui, ul, rg = P.swapped_info(g["y"][i], g["labels"][i].astype(np.int64))
        assert np.array_equal(ui, g["upd_in"][i]) and np.array_equal(ul, g["upd_lab"][i])
        assert rg.dtype == np.int32 and np.array_equal(rg, np.unpackbits(g["redG"][i], axis=1)[:, :128])
        for o in (1, 2):
            ok, T, phase = C.convention_osd_main((ui, ul, rg, C.generate_teps(o), C.query_boundary(o)))
            assert (ok, T, phase) == (bool(g[f"ok{o}"][i]), [1, 65, 2081][o], int(g[f"phase{o}"][i]))
        # PB copy: 6-tuple with the original inputs weighting the discrepancy
        ok6, _, _ = C.convention_osd_main((ui, ui, ul, rg, C.generate_teps(1), C.query_boundary(1)))
        assert ok6 == bool(g["ok1"][i])
    res = C.convention_osd_batch(g["y"], g["labels"], 2)
    assert np.array_equal(res["correct"], g["ok2"]) and np.array_equal(res["phase"], g["phase2"])


def test_fs_kernel_matches_reference_policy(handle, golden_dir):
    from short_ldpc_decoding_osd_b200 import fs_testing as F

    g = load(golden_dir, "fs_ref_shim.npz")
    G = load(golden_dir, "code_ref.npz")["G"].astype(np.int64)
    for order in (1, 2, 3):  # 3 = the reference's default order_limit; its golden covers the first 12 frames
        GL.set_map("order_limit", order)
        n = len(g[f"success{order}"])
        if order < 3:
            seq = F.generate_sequential_teps(64, order)
            assert np.array_equal(np.concatenate(seq, 0), np.unpackbits(g[f"seq{order}"], axis=1)[:, :64])
        res = F.fs_osd_batch(g["y"][:n], g["labels"][:n], order, float(g["beta"]))
        assert np.array_equal(res["correct"].astype(int), g[f"success{order}"])   # the reference's own S/F per frame
        assert np.array_equal(res["num_teps"], g[f"num_teps{order}"])             # and its TEP counts
        for i in range(n if order < 3 else 4):
            ref = OO.fs_frame(g["y"][i], G, order, 6.5, 30, 0.1)
            assert np.array_equal(res["codeword"][i], ref["codeword"])
            assert (int(res["best_tep"][i]), int(res["num_teps"][i]), int(res["stop_kind"][i])) == (ref["best_tep"], ref["num_teps"], ref["stop_kind"])
    GL.set_map("order_limit", 2)


def test_fs_policy_bit_exact_on_more_frames(handle, code):
    from oracle import philox_oracle as PO
    from short_ldpc_decoding_osd_b200 import fs_testing as F

    y, cw, _ = PO.gen_frames(77, 0, 1500, 2.0, code.G)
    syn = nms_gpu(handle, y, 12, traj=False)["syndrome_nz"]
    yf, cf = y[syn][:120], cw[syn][:120]
    for order, tau_e, tau_psc, beta in [(2, 6.5, 30, 0.1), (1, 6.5, 30, 0.1), (2, 9.5, 24, 0.0), (3, 6.5, 30, 0.3), (0, 6.5, 30, 0.1)]:
        n = len(yf) if order < 3 else 12
        res = F.fs_osd_batch(yf[:n], cf[:n], order, beta, tau_e, tau_psc)
        for i in range(n):
            ref = OO.fs_frame(yf[i], code.G, order, tau_e, tau_psc, beta)
            assert np.array_equal(res["codeword"][i], ref["codeword"]), (order, i)
            assert (int(res["best_tep"][i]), int(res["num_teps"][i]), int(res["stop_kind"][i])) == (ref["best_tep"], ref["num_teps"], ref["stop_kind"]), (order, i)
    assert len(set(res["stop_kind"])) >= 1


def test_dl_sliding_osd_matches_reference(handle, code, golden_dir):
    from short_ldpc_decoding_osd_b200 import nn_testing as NT
    from short_ldpc_decoding_osd_b200 import ordered_statistics_decoding as OSD

    g = load(golden_dir, "dl_ref_shim.npz")
    for k, v in dict(threshold_sum=2, segment_num=6, soft_margin=0.9, decoding_length=30, sliding_win_width=5).items():
        GL.set_map(k, v)
    osd = OSD.osd(code)
    tep_info = NT.generate_teps(osd, [list(p) for p in g["path"]])
    assert [b.shape[0] for b in tep_info[0]] == list(g["block_sizes"])
    assert np.array_equal(np.concatenate(tep_info[0], 0), np.unpackbits(g["blocks"], axis=1)[:, :64])
    W, V, b0 = g["W"], g["V"], float(g["fcn_bias0"])

    def fcn(x):
        o = (np.asarray(x, dtype=np.float32) @ W) @ V
        o[..., 0] += b0
        e = np.exp(o - o.max(axis=-1, keepdims=True))
        return e / e.sum(axis=-1, keepdims=True)

    B = len(g["y"])
    input_list = g["traj"].reshape(-1, 128)
    bm, tq, ex, pm = osd.block_minima(input_list, g["new_inputs"], g["labels"], tep_info)
    for i in range(B):
        # same MRB as the reference's H-based elimination, least reliable first
        assert np.array_equal(pm[i][:64][::-1], g["lri"][i][g["upd_idx"][i][64:]])
        mins = g["block_mins_fp32"][i]
        k = int(np.sum(~np.isnan(mins)))
        np.testing.assert_allclose(bm[i][:k].astype(np.float64) * 2.0 ** (int(ex[i]) - 54), mins[:k], rtol=2e-6)
    for i in range(B):
        s, f, w, c = osd.sliding_osd(fcn, input_list[13 * i:13 * i + 13], g["new_inputs"][i:i + 1], g["labels"][i:i + 1], tep_info)
        rs, rf, rw, rc = (int(v) for v in g["per_frame"][i])
        assert (w, c) == (rw, rc)            # windows visited and TEP complexity: identical policy trace
        if rs:                                # fp32 equality may report a false failure (frame 7, see the oracle test)
            assert s == 1
    # the same policy on the GPU: the fixture's classifier as Predict_outlier_light kernels (bias folded into W2 via k? no:
    # the fixture adds a constant to logit 0; a constant column cannot be expressed, so compare GPU vs host on a pure-kernel net)
    from short_ldpc_decoding_osd_b200 import nn_net as NN

    net = NN.Predict_outlier_light(5, W1=np.eye(6, dtype=np.float32), W2=np.array([[0, -0.5], [0, 0.5], [0, 0], [0, 0], [0, 0], [0, 0.15]], np.float32))
    host = [osd.__class__.sliding_osd(osd, (lambda x: net(x)), input_list[13 * i:13 * i + 13], g["new_inputs"][i:i + 1], g["labels"][i:i + 1], tep_info) for i in range(B)]
    gpu = osd.sliding_osd(net, input_list, g["new_inputs"], g["labels"], tep_info)
    assert gpu == tuple(int(sum(x[j] for x in host)) for j in range(4))
    assert len({x[2] for x in host}) > 1  # the policy actually stops at different depths
    # DIA FIR drop-in reproduces the fixture's ordering metric
    from short_ldpc_decoding_osd_b200 import nn_net

    nn = nn_net.conv_bitwise()
    nn.set_taps(g["taps"])
    squashed, inputs0, labels0 = nn.preprocessing_inputs((input_list, np.repeat(g["labels"], 13, axis=0)))
    assert squashed.shape == (B * 128, 13, 1) and np.array_equal(inputs0, g["y"]) and np.array_equal(labels0, g["labels"])
    np.testing.assert_allclose(nn(squashed), g["new_inputs"], rtol=1e-5, atol=1e-5)


def test_pb_kernel_matches_reference_policy(handle, golden_dir):
    """PB-OSD: per-frame S/F, TEPs visited and both improvement counters of the reference's pb_osd."""
    from oracle import pb_oracle as PB
    from short_ldpc_decoding_osd_b200 import pb_testing as P

    g = load(golden_dir, "pb_ref_shim.npz")
    G = load(golden_dir, "code_ref.npz")["G"].astype(np.int64)
    for order, snr, tag in ((1, 2.5, "o1_snr25"), (2, 2.5, "o2_snr25"), (2, 3.5, "o2_snr35"), (3, 2.5, "o3_snr25")):
        res = P.pb_osd_batch(g["y"], g["labels"], snr, order)
        assert np.array_equal(res["correct"].astype(int), g[f"success_{tag}"])
        assert np.array_equal(res["num_teps"], g[f"num_teps_{tag}"])
        assert np.array_equal(res["suc1"], g[f"suc1_{tag}"]) and np.array_equal(res["suc2"], g[f"suc2_{tag}"])
        assert np.array_equal(res["list_cmp"], g[f"list_cmp_{tag}"])
        for i in range(0, len(g["y"]), 4):
            ref = PB.pb_frame(g["y"][i], G, snr, order)
            assert np.array_equal(res["codeword"][i], ref["codeword"])


def test_pb_policy_on_more_frames(handle, code):
    from oracle import pb_oracle as PB
    from oracle import philox_oracle as PO
    from short_ldpc_decoding_osd_b200 import pb_testing as P

    y, cw, _ = PO.gen_frames(123, 0, 1500, 2.5, code.G)
    syn = nms_gpu(handle, y, 12, traj=False)["syndrome_nz"]
    yf, cf = y[syn][:150], cw[syn][:150]
    for order, snr in ((2, 2.5), (1, 3.0), (0, 2.5), (3, 3.0)):
        res = P.pb_osd_batch(yf, cf, snr, order)
        mism = 0
        for i in range(len(yf)):
            ref = PB.pb_frame(yf[i], code.G, snr, order)
            same = np.array_equal(res["codeword"][i], ref["codeword"]) and (int(res["num_teps"][i]), int(res["suc1"][i]), int(res["suc2"][i]), int(res["list_cmp"][i])) == (ref["num_teps"], ref["suc1"], ref["suc2"], ref["list_cmp"])
            mism += not same
        # probabilities are fp32 with expf/pow of two different libms: a stop decision may flip when a probability
        # sits within an ulp of its threshold
        assert mism <= 1, (order, snr, mism)


def test_driver_chain_nms_to_retest_file_to_osd(handle, code, tmp_path, monkeypatch):
    """The reference's file-coupled chain end to end with the drop-ins: ldpc_128_testing -> retest TFRecord
    (13 rows per failure) -> fs_osd / pb_osd drivers reading it in batches of 13 (globalmap.data_setting)."""
    from short_ldpc_decoding_osd_b200 import fs_testing, ldpc_128_testing, pb_testing, read_TFdata

    monkeypatch.chdir(tmp_path)
    argv = "python 2.5 2.5 1 500 4 12 CCSDS_ldpc_n128_k64.alist NMS-1".split()
    fer_list = ldpc_128_testing.main(argv, data_root=str(tmp_path), frames_if_missing=2000)
    assert len(fer_list) == 1 and 0.15 < fer_list[0][1] < 0.32
    retest = tmp_path / "Testing_data_gen_128" / "data" / "snr2.5-2.5dB" / "NMS-1" / "12th" / "2.5dB" / "ldpc-nonzero-retest.tfrecord"
    assert retest.exists()
    ds = read_TFdata.data_handler(128, str(retest), 13)
    batches = list(ds.as_numpy_iterator())
    n_fail = len(batches)
    assert all(b[0].shape == (13, 128) for b in batches) and abs(n_fail - fer_list[0][1] * 2000) <= 2
    log = open(tmp_path / "log" / "FER-NMS-1-12th.txt").read()
    assert "FER 0." in log and "CE list" in log
    GL.set_map("order_limit", 1)
    for k, v in dict(convention_osd=False, miracle_view=False, fs_osd=True, pb_osd=True, termination_num_threshlod=100000).items():
        GL.set_map(k, v)
    s_fs = fs_testing.fs_osd(2.5, 0.1, ds)
    s_pb = pb_testing.pb_osd(2.5, ds)
    GL.set_map("convention_osd", True)
    s_cv = fs_testing.fs_osd(2.5, 0.1, ds)
    GL.set_map("convention_osd", False)
    assert s_fs["S"] + s_fs["F"] == s_pb["S"] + s_pb["F"] == s_cv["S"] + s_cv["F"] == n_fail
    assert s_cv["F"] <= s_fs["F"] and s_cv["F"] <= s_pb["F"]  # the exhaustive sweep lower-bounds both policies
    assert os.path.exists(s_fs["log"]) and os.path.exists(s_pb["log"])
    GL.set_map("order_limit", 2)
