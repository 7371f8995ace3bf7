"""CPU tests of the host-side logic of the drop-in modules (no GPU): configuration mirror, TEP packing,
DIA folding, TFRecord IO, FER confidence intervals, frame sharding and the counter all-reduce (gloo, 2 ranks)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import osd_oracle as OO
from short_ldpc_decoding_osd_b200 import convention_osd as C
from short_ldpc_decoding_osd_b200 import globalmap as GL
from short_ldpc_decoding_osd_b200 import nn_net, nn_testing, read_TFdata, simulate
from short_ldpc_decoding_osd_b200 import ordered_statistics_decoding as OSD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_globalmap_mirror(capsys):
    GL.global_setting("python 2.0 3.0 6 1 12 CCSDS_ldpc_n128_k64.alist NMS-1".split())
    assert GL.get_map("order_limit") == 3 and GL.get_map("termination_num_threshlod") == 100
    assert GL.get_map("tau_psc") == 30 and GL.get_map("d_min") == 14 and GL.get_map("soft_margin") == 0.9
    assert GL.get_map("code_parameters").k == 64
    assert GL.get_map("no-such-key") is None and "non-existence" in capsys.readouterr().out
    sizes, bnd = GL.secure_segment_threshold()
    assert list(sizes) == [1, 4, 8, 12, 16, 23] and list(bnd) == [0, 1, 5, 13, 25, 41, 64]
    assert C.query_boundary(3) == [1, 65, 2081, 43745]


def test_tep_packing_roundtrip():
    for teps in (OO.generate_teps_conv(2), OO.generate_teps_fs(2)):
        words = OO.pack_teps(teps)
        m = C.unpack_tep_words(words)
        assert m.shape == (2081, 64) and np.array_equal(C.pack_tep_matrix(m), words)
        assert all(tuple(np.flatnonzero(m[i])) == tuple(sorted(teps[i])) for i in range(0, 2081, 97))
    with pytest.raises(ValueError):
        C.pack_tep_matrix(np.ones((1, 64), dtype=int))
    dl = np.zeros((2, 64), dtype=int)
    dl[1, [0, 5]] = 1
    w = OSD.pack_dl_teps(dl)
    assert w[0] == 0xFFFFFFFF and (w[1] & 0xFF, (w[1] >> 8) & 0xFF) == (58, 63)


def test_convention_and_segment_paths():
    GL.set_map("threshold_sum", 3)
    GL.set_map("segment_num", 6)
    GL.set_map("decoding_length", 30)
    GL.set_map("code_parameters", GL.get_map("code_parameters") or __import__("short_ldpc_decoding_osd_b200.fill_matrix_info", fromlist=["Code"]).Code())
    path, nn_type = nn_testing.query_convention_path()
    assert nn_type == "benchmark" and path[0] == [0, 0, 0] and len(path) == len({tuple(p) for p in path}) == 20
    seg = nn_testing.convention_segment_path()
    # 84 compositions of <= 3 over 6 segments, minus the 7 that need >= 2 positions from the 1-wide first segment;
    # the 77 non-empty blocks cover all 43,745 TEPs (SURVEY 8a a11''')
    assert len(seg) == 77 and seg[0] == [0] * 6
    osd = OSD.osd(GL.get_map("code_parameters"))
    blocks, acc = nn_testing.generate_teps(osd, seg)
    assert acc[-1] == 43745 and len(nn_testing.filter_order_patterns(seg)) == 30


def test_dia_cnn_folds_to_fir():
    rng = np.random.default_rng(2)
    k1, k2, k3 = rng.normal(size=(3, 1, 8)), rng.normal(size=(3, 8, 4)), rng.normal(size=(3, 4, 2))
    dw, db = rng.normal(size=(14, 1)), rng.normal(size=1)
    taps, bias = nn_net.fold_conv_bitwise(k1, k2, k3, dw, db)
    x = rng.normal(size=13)

    def conv(v, k):
        return np.stack([np.einsum("dc,dco->o", v[t:t + 3], k) for t in range(v.shape[0] - 2)])

    direct = conv(conv(conv(x.reshape(-1, 1), k1), k2), k3).reshape(-1) @ dw.reshape(-1) + db[0]
    assert taps.shape == (13,) and np.isclose(taps.astype(np.float64) @ x + bias, direct, rtol=1e-5)
    p = nn_net.Predict_outlier_light(5)(np.arange(6, dtype=np.float32).reshape(1, 6))
    assert p.shape == (1, 2) and np.isclose(p.sum(), 1.0)


def test_tfrecord_roundtrip_and_corruption(tmp_path):
    assert read_TFdata.crc32c(b"123456789") == 0xE3069283  # CRC-32C check value
    rng = np.random.default_rng(0)
    f = rng.normal(size=(27, 128)).astype(np.float32)
    lab = rng.integers(0, 2, (27, 128))
    p = str(tmp_path / "ldpc-nonzero-retest.tfrecord")
    read_TFdata.make_tfrecord((f, lab), p)
    batches = list(read_TFdata.data_handler(128, p, 13).as_numpy_iterator())
    assert [b[0].shape[0] for b in batches] == [13, 13, 1] and batches[0][1].dtype == np.int64
    assert np.array_equal(np.concatenate([b[0] for b in batches]), f)
    assert np.array_equal(np.concatenate([b[1] for b in batches]), lab)
    assert list(batches[0][2][:2]) == [128, 128]
    raw = bytearray(open(p, "rb").read())
    raw[40] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    with pytest.raises(IOError):
        list(read_TFdata.data_handler(128, p, 13).as_numpy_iterator())
    ex = read_TFdata.parse_example(read_TFdata.serialize_example(np.array([1.5, -2.0], np.float32), np.array([1, -3])))
    assert list(ex["feature"]) == [1.5, -2.0] and list(ex["label"]) == [1, -3] and list(ex["shape"]) == [2]
    empty = str(tmp_path / "empty.tfrecord")
    open(empty, "wb").close()
    assert list(read_TFdata.data_handler(128, empty, 4).as_numpy_iterator()) == []


def test_confidence_intervals_and_shards():
    lo, hi = simulate.wilson_interval(0, 1000)
    assert lo < 1e-12 and 0 < hi < 0.005
    lo, hi = simulate.wilson_interval(230, 1000)
    assert lo < 0.23 < hi and hi - lo < 0.06
    lo2, hi2 = simulate.clopper_pearson(230, 1000)
    assert lo2 < 0.23 < hi2 and abs(lo2 - lo) < 0.005
    assert simulate.wilson_interval(0, 0) == (0.0, 1.0)
    # contiguous, disjoint, exhaustive frame ranges
    for total, world in ((1000, 1), (1000, 3), (7, 8), (1 << 24, 8)):
        shards = [simulate.shard_range(total, r, world) for r in range(world)]
        assert shards[0][0] == 0 and shards[-1][1] == total
        assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in shards) - min(b - a for a, b in shards) <= 1
    c = simulate.Tallies(np.array([1000, 230, 2500, 229, 1, 12000, 229, 20, 300, 21, 310, 476549, 10, 120, 79, 0], dtype=np.uint64))
    assert c.fer_nms == 0.23 and c.fer_final == 0.021 and c.osd_frames == 229 and c.ber_nms == 2500 / 128000
    assert "FER 0.2300" in c.reference_log_line()


def test_counter_allreduce_two_ranks_gloo(tmp_path):
    """The only collective of the path: a sum all-reduce of the 16 uint64 counters (world size 2, gloo)."""
    script = tmp_path / "w.py"
    script.write_text(
        "import sys, numpy as np, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from short_ldpc_decoding_osd_b200 import simulate\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "a, b = simulate.shard_range(1001, r, w)\n"
        "local = np.zeros(16, dtype=np.uint64); local[0] = b - a; local[1] = 7 * (r + 1); local[11] = (1 << 40) + r\n"
        "tot = simulate.allreduce_counters(local)\n"
        "assert tot[0] == 1001 and tot[1] == 21 and tot[11] == (1 << 41) + 1, tot\n"
        "stop = simulate.should_stop(tot, max_frame_errors=20, counter='nms_frame_err')\n"
        "assert stop is True\n"
        "sys.stdout.write(f'rank{r}ok\\n')\n"
    )
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_exported_weights_round_trip(tmp_path):
    """weights.npz / decoding_path.json as scripts/export_tf_weights.py writes them."""
    import json

    from short_ldpc_decoding_osd_b200 import nn_net, weights

    rng = np.random.default_rng(0)
    w = {"nms_check": np.array([-0.048], np.float32), "k1": rng.normal(size=(3, 1, 8)), "k2": rng.normal(size=(3, 8, 4)),
         "k3": rng.normal(size=(3, 4, 2)), "dense_w": rng.normal(size=(14, 1)), "dense_b": np.array([0.1]),
         "fcn1": rng.normal(size=(6, 6)), "fcn2": rng.normal(size=(6, 2))}
    p = tmp_path / "weights.npz"
    np.savez(p, **w)
    got = weights.load_npz(str(p))
    assert abs(weights.alpha_of(got) - 0.66943514) < 1e-6
    taps, bias = nn_net.fold_conv_bitwise(got["k1"], got["k2"], got["k3"], got["dense_w"], got["dense_b"])
    # the folded FIR equals the three linear convolutions + dense layer on a random trajectory
    x = rng.normal(size=13)
    h = x.reshape(-1, 1)
    for k in (w["k1"], w["k2"], w["k3"]):
        h = np.stack([np.einsum("dc,dco->o", h[t:t + 3], k) for t in range(h.shape[0] - 2)])
    want = float(h.reshape(-1) @ w["dense_w"].reshape(-1) + 0.1)
    assert abs(float(taps.astype(np.float64) @ x + bias) - want) < 1e-4
    fcn = weights.make_fcn(got)
    pr = fcn(rng.normal(size=(3, 6)).astype(np.float32))
    assert pr.shape == (3, 2) and np.allclose(pr.sum(1), 1.0, atol=1e-6)
    np.savez(p, k1=np.zeros((3, 2, 8)))
    with pytest.raises(ValueError):
        weights.load_npz(str(p))
    q = tmp_path / "path.json"
    q.write_text(json.dumps({"decoding_path": [[0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0]], "counts": [9, 5, 2]}))
    assert weights.load_decoding_path(str(q))[1] == [1, 0, 0, 0, 0, 0]


def test_generator_replays_the_reference_numpy_stream(golden_dir):
    """north_star: the frame generator "can replay the reference's noise seeds".  With replay_numpy_seed the drop-in
    returns exactly what np.random.seed(s) + the reference's testing_data_generating return (fixture: the reference's
    own function run with seed 0, 4000 frames, 2.5 dB -- oracle/ref_runner.py run_gen)."""
    import os

    from short_ldpc_decoding_osd_b200 import data_generating as DG
    from short_ldpc_decoding_osd_b200.fill_matrix_info import Code

    g = np.load(os.path.join(golden_dir, "gen_ref_stats.npz"))
    code = Code()
    y, lab = DG.testing_data_generating(code, float(g["snr"]), int(g["n"]), replay_numpy_seed=0)
    assert y.dtype == np.float64 and y.shape == (int(g["n"]), 128)
    assert np.array_equal(y[:8].astype(np.float32), g["y_head"])
    assert np.array_equal(lab[:8].astype(np.uint8), g["labels_head"])
    z = np.where(lab == 0, y, -y) - 1.0
    assert abs(float(z.std()) - float(g["sigma_hat"])) < 1e-6 and abs(float(lab.mean()) - float(g["ones_frac"])) < 1e-12  # the fixture took its moments on the float32 copy


def test_retest_buffer_row_sequences_behave_like_the_reference_lists():
    """ms_test._Rows: the (buffer_inputs, buffer_labels) Decoding_model returns are row sequences over one array each;
    the reference's drivers index, slice, iterate, len() and concatenate them (ldpc_128_testing.py:123-125,
    ms_test.py:66-70)."""
    from short_ldpc_decoding_osd_b200 import ms_test as M

    base = np.arange(6 * 128, dtype=np.float32).reshape(6, 128)
    rows = M._Rows(base)
    assert len(rows) == 6 and np.array_equal(rows[2], base[2]) and np.array_equal(rows[-1], base[5])
    assert [r[0] for r in rows] == [float(base[i, 0]) for i in range(6)]
    assert np.array_equal(np.stack(rows[1:3]), base[1:3]) and np.array_equal(np.asarray(rows), base)
    lab = np.arange(3 * 128).reshape(3, 128)
    rep = M._Rows(lab, np.repeat(np.array([2, 0]), 13))          # label of failure i repeated 13 times
    assert len(rep) == 26 and np.array_equal(rep[0], lab[2]) and np.array_equal(rep[13], lab[0])
    assert np.array_equal(rep.array(), lab[np.repeat(np.array([2, 0]), 13)])
    assert len(rows + rep) == 32 and len([1] + rows) == 7        # list concatenation both ways
    model = M.Decoding_model.__new__(M.Decoding_model)
    flat_i, flat_l = model.postprocess_failure_cases(([rows, rows], [rep, rep]))
    assert len(flat_i) == 12 and len(flat_l) == 52 and np.array_equal(flat_i.array()[6:], base)
    # the reference's list-of-lists form still works
    flat_i, flat_l = model.postprocess_failure_cases(([[base[0], base[1]]], [[lab[0], lab[1]]]))
    assert len(flat_i) == 2 and np.array_equal(flat_i[1], base[1])
