"""-m gpu tests of the fused device pipeline, tallies, compaction and the host-buffer entry points."""
import numpy as np
import pytest
import torch

from oracle import nms_oracle as NO
from oracle import osd_oracle as OO
from oracle import philox_oracle as PO
from short_ldpc_decoding_osd_b200 import _lib
from tests.gpu_util import dev, empty, nms_gpu, osd_gpu, sync

pytestmark = pytest.mark.gpu
ALPHA = float(NO.softplus(-0.048))
CN = {n: i for i, n in enumerate(_lib.COUNTER_NAMES)}


def expected_counters(cw, nms_hard, syn, iters, final, best_tep, n_teps):
    c = np.zeros(16, dtype=np.uint64)
    d_n = (nms_hard != cw).sum(1)
    d_f = (final != cw).sum(1)
    c[CN["frames"]] = len(cw)
    c[CN["nms_frame_err"]] = (d_n > 0).sum()
    c[CN["nms_bit_err"]] = d_n.sum()
    c[CN["nms_detected"]] = syn.sum()
    c[CN["nms_undetected"]] = ((~syn) & (d_n > 0)).sum()
    c[CN["nms_iters"]] = iters.sum()
    c[CN["osd_frames"]] = syn.sum()
    c[CN["osd_frame_err"]] = (syn & (d_f > 0)).sum()
    c[CN["osd_bit_err"]] = d_f[syn].sum()
    c[CN["final_frame_err"]] = (d_f > 0).sum()
    c[CN["final_bit_err"]] = d_f.sum()
    c[CN["teps"]] = syn.sum() * n_teps
    ok = syn & (d_f == 0)
    for w, (lo, hi) in enumerate([(0, 1), (1, 65), (65, 2081), (2081, 43745)]):
        c[CN["phase0"] + w] = (ok & (best_tep >= lo) & (best_tep < hi)).sum()
    return c


@pytest.mark.parametrize("order,B", [(1, 3000), (2, 1500), (-1, 500)])
def test_decode_pipeline_equals_composition(handle, code, order, B):
    y, cw, _ = PO.gen_frames(4, 0, B, 2.5, code.G)
    yd = dev(y)
    truth = dev(_lib.pack_bits(cw).view(np.int32))
    bits = empty((B, 4), torch.int32)
    syn = empty((B,), torch.uint8)
    bt = empty((B,), torch.int32)
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda:0")
    handle.call("ldpcb_decode", yd, B, 12, ALPHA, 1.0, 1.0, 0, order, 0, bits, syn, bt, truth, cnt, None)
    sync()
    # composition of the two stand-alone kernels
    nms = nms_gpu(handle, y, 12, ALPHA, traj=False)
    final = nms["hard"].copy()
    best = np.full(B, -1, dtype=np.int64)
    fails = np.flatnonzero(nms["syndrome_nz"])
    if order >= 0 and len(fails):
        osd = osd_gpu(handle, y[fails], order=order)
        final[fails] = osd["codeword"]
        best[fails] = osd["best_tep"]
    got_bits = _lib.unpack_bits(bits.cpu().numpy().view(np.uint32))
    assert np.array_equal(syn.cpu().numpy().astype(bool), nms["syndrome_nz"])
    assert np.array_equal(got_bits, final)
    assert np.array_equal(bt.cpu().numpy(), best)
    n_teps = handle.tep_count(order) if order >= 0 else 0
    synm = nms["syndrome_nz"] if order >= 0 else np.zeros(B, dtype=bool)
    exp = expected_counters(cw, nms["hard"], nms["syndrome_nz"], nms["iters_used"], final, best, n_teps)
    if order < 0:
        for k in ("osd_frames", "osd_frame_err", "osd_bit_err", "teps", "phase0", "phase1", "phase2", "phase3"):
            exp[CN[k]] = 0
    assert np.array_equal(cnt.cpu().numpy().view(np.uint64), exp)
    # OSD can only lower the frame error count here
    assert exp[CN["final_frame_err"]] <= exp[CN["nms_frame_err"]]


def test_simulate_equals_decode_of_generated_frames(handle, code):
    B, seed, first = 5000, 17, 123456
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda:0")
    handle.call("ldpcb_simulate", seed, first, B, 3.0, 12, ALPHA, 1.0, 1.0, 0, 1, 0, cnt, None)
    yd = empty((B, 128), torch.float32)
    truth = empty((B, 4), torch.int32)
    handle.call("ldpcb_gen_frames", seed, first, B, 3.0, yd, truth, None)
    bits = empty((B, 4), torch.int32)
    cnt2 = torch.zeros(16, dtype=torch.int64, device="cuda:0")
    handle.call("ldpcb_decode", yd, B, 12, ALPHA, 1.0, 1.0, 0, 1, 0, bits, None, None, truth, cnt2, None)
    sync()
    a, b = cnt.cpu().numpy(), cnt2.cpu().numpy()
    assert a[CN["frames"]] == B
    # the phase histogram needs best_tep, which ldpcb_decode was not given here
    assert np.array_equal(a[:12], b[:12])
    # sharding: two halves accumulate to the same counters
    cnt3 = torch.zeros(16, dtype=torch.int64, device="cuda:0")
    handle.call("ldpcb_simulate", seed, first, 2000, 3.0, 12, ALPHA, 1.0, 1.0, 0, 1, 0, cnt3, None)
    handle.call("ldpcb_simulate", seed, first + 2000, 3000, 3.0, 12, ALPHA, 1.0, 1.0, 0, 1, 0, cnt3, None)
    sync()
    assert np.array_equal(a, cnt3.cpu().numpy())


def test_select_flagged_and_gather(handle):
    rng = np.random.default_rng(3)
    for B in (1, 5, 2048, 2049, 70001):
        flags = (rng.random(B) < 0.2).astype(np.uint8)
        fd = dev(flags)
        idx = empty((B,), torch.int32)
        cnt = empty((1,), torch.int32)
        handle.call("ldpcb_select_flagged", fd, B, idx, cnt, None)
        sync()
        n = int(cnt.item())
        want = np.flatnonzero(flags)
        assert n == len(want)
        assert np.array_equal(idx[:n].cpu().numpy(), want)
        src = torch.arange(B * 8, dtype=torch.float32, device="cuda:0").reshape(B, 8)
        dst = torch.zeros((B, 8), dtype=torch.float32, device="cuda:0")
        handle.call("ldpcb_gather_rows", src, idx, cnt, B, 8, dst, None)
        sync()
        assert torch.equal(dst[:n], src[torch.from_numpy(want).to("cuda:0")])
    cnt = empty((1,), torch.int32)
    handle.call("ldpcb_select_flagged", None, 0, None, cnt, None)
    sync()
    assert int(cnt.item()) == 0


def test_dia_fir(handle):
    rng = np.random.default_rng(1)
    traj = rng.normal(size=(37, 13, 128)).astype(np.float32)
    taps = rng.normal(size=13).astype(np.float32)
    out = empty((37, 128), torch.float32)
    handle.call("ldpcb_dia_fir", dev(traj), 37, 13, taps, 0.25, out, None)
    sync()
    ref = np.einsum("bij,i->bj", traj.astype(np.float64), taps.astype(np.float64)) + 0.25
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)



def test_nms_decode_fir_equals_trajectory_then_fir(handle, code):
    """ldpcb_nms_decode_fir (DIA FIR fused into the decoder) is bit-identical to ldpcb_nms_decode(soft_traj) followed by
    ldpcb_dia_fir, and its hard decisions / syndrome flags equal the plain decode."""
    B = 3001
    y = empty((B, 128), torch.float32)
    tr = empty((B, 4), torch.int32)
    handle.call("ldpcb_gen_frames", 55, 0, B, 2.0, y, tr, None)
    rng = np.random.default_rng(8)
    taps = (np.full(13, 1 / 13) + 0.05 * rng.normal(size=13)).astype(np.float32)
    bits = empty((B, 4), torch.int32)
    syn = empty((B,), torch.uint8)
    traj = empty((B, 13, 128), torch.float32)
    handle.call("ldpcb_nms_decode", y, B, 12, ALPHA, 1.0, 1.0, 0, bits, None, syn, traj, None)
    want = empty((B, 128), torch.float32)
    handle.call("ldpcb_dia_fir", traj, B, 13, taps, 0.07, want, None)
    bits2 = empty((B, 4), torch.int32)
    syn2 = empty((B,), torch.uint8)
    got = empty((B, 128), torch.float32)
    handle.call("ldpcb_nms_decode_fir", y, B, 12, ALPHA, 1.0, 1.0, taps, 0.07, bits2, syn2, got, None)
    sync()
    assert torch.equal(got, want)
    assert torch.equal(bits, bits2) and torch.equal(syn, syn2)


def test_host_entry_points_equal_device_ones(handle, code):
    B = 70000  # more than one host chunk
    y, cw, _ = PO.gen_frames(8, 0, B, 2.5, code.G)
    yp = _lib.pinned_empty((B, 128), np.float32)
    yp[:] = y
    bits = np.empty((B, 4), np.uint32)
    it = np.empty(B, np.uint8)
    syn = np.empty(B, np.uint8)
    handle.call("ldpcb_nms_decode_host", yp, B, 12, ALPHA, 1.0, 1.0, 0, bits, it, syn, None)
    dev_res = nms_gpu(handle, y, 12, ALPHA, traj=False)
    assert np.array_equal(_lib.unpack_bits(bits), dev_res["hard"])
    assert np.array_equal(syn.astype(bool), dev_res["syndrome_nz"])
    # with the trajectory, small batch, pageable memory
    tr = np.empty((100, 13, 128), np.float32)
    b2 = np.empty((100, 4), np.uint32)
    handle.call("ldpcb_nms_decode_host", np.ascontiguousarray(y[:100]), 100, 12, ALPHA, 1.0, 1.0, 0, b2, None, None, tr)
    ref = nms_gpu(handle, y[:100], 12, ALPHA)
    assert np.array_equal(tr, ref["traj"])
    # whole pipeline from host buffers
    truth = _lib.pack_bits(cw)
    fb = np.empty((B, 4), np.uint32)
    bt = np.empty(B, np.int32)
    cnt = np.zeros(16, np.uint64)
    handle.call("ldpcb_decode_host", yp, B, 12, ALPHA, 1.0, 1.0, 0, 1, 0, fb, syn, bt, truth, cnt)
    yd = dev(y)
    bits_d = empty((B, 4), torch.int32)
    cnt_d = torch.zeros(16, dtype=torch.int64, device="cuda:0")
    bt_d = empty((B,), torch.int32)
    handle.call("ldpcb_decode", yd, B, 12, ALPHA, 1.0, 1.0, 0, 1, 0, bits_d, None, bt_d, dev(truth.view(np.int32)), cnt_d, None)
    sync()
    assert np.array_equal(fb.view(np.int32), bits_d.cpu().numpy())
    assert np.array_equal(bt, bt_d.cpu().numpy())
    assert np.array_equal(cnt, cnt_d.cpu().numpy().view(np.uint64))
    # OSD from host buffers
    fails = np.flatnonzero(syn)[:500]
    cwb = np.empty((len(fails), 4), np.uint32)
    bt2 = np.empty(len(fails), np.int32)
    yo = np.ascontiguousarray(y[fails])
    handle.call("ldpcb_osd_decode_host", yo, yo, len(fails), 1, 0, 0, cwb, bt2, None, None, None, None)
    assert np.array_equal(cwb, fb[fails])
    assert np.array_equal(bt2, bt[fails])


def test_large_batch_properties(handle, code):
    """Full-size properties: every OSD output is a codeword, OSD never changes converged frames,
    counters add up, FER in the range the CPU oracle gives at 2.5 dB."""
    B = 1 << 20
    yd = empty((B, 128), torch.float32)
    truth = empty((B, 4), torch.int32)
    handle.call("ldpcb_gen_frames", 2024, 0, B, 2.5, yd, truth, None)
    bits = empty((B, 4), torch.int32)
    syn = empty((B,), torch.uint8)
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda:0")
    handle.call("ldpcb_decode", yd, B, 12, ALPHA, 1.0, 1.0, 0, 2, 0, bits, syn, None, truth, cnt, None)
    sync()
    c = cnt.cpu().numpy()
    assert c[CN["frames"]] == B
    fer_nms = c[CN["nms_frame_err"]] / B
    assert 0.19 < fer_nms < 0.25
    assert c[CN["final_frame_err"]] == c[CN["nms_undetected"]] + c[CN["osd_frame_err"]]
    assert c[CN["final_frame_err"]] / B < 0.03
    # all OSD outputs satisfy H c = 0: syndrome through the packed check masks
    Hm = torch.from_numpy(_lib.pack_bits(code.H).view(np.int32)).to("cuda:0")  # [64,4]
    sel = bits[syn.bool()]
    par = torch.zeros((sel.shape[0], 64), dtype=torch.int32, device="cuda:0")
    for w in range(4):
        x = sel[:, w:w + 1] & Hm[None, :, w]
        for s in (16, 8, 4, 2, 1):
            x = x ^ (x >> s)
        par ^= x & 1
    assert int(par.sum()) == 0


def test_dl_device_pipeline_equals_module_chain(handle, code):
    """simulate.run_point_dl (all on device) against the drop-in module chain Decoding_model -> conv_bitwise ->
    osd.sliding_osd on the same frames."""
    from short_ldpc_decoding_osd_b200 import globalmap as GL
    from short_ldpc_decoding_osd_b200 import ms_test, nn_net, nn_testing, simulate
    from short_ldpc_decoding_osd_b200 import ordered_statistics_decoding as OSD

    for k, v in dict(code_parameters=code, selected_decoder_type="NMS-1", num_iterations=12, threshold_sum=2, segment_num=6, soft_margin=0.9,
                     decoding_length=30, sliding_win_width=5).items():
        GL.set_map(k, v)
    osd = OSD.osd(code)
    path = nn_testing.filter_order_patterns(nn_testing.convention_segment_path())
    tep_info = nn_testing.generate_teps(osd, path)
    rng = np.random.default_rng(3)
    taps = (np.full(13, 1 / 13) + 0.03 * rng.normal(size=13)).astype(np.float32)
    net = nn_net.Predict_outlier_light(5, W1=np.eye(6, dtype=np.float32), W2=np.array([[0, -0.5], [0, 0.5], [0, 0], [0, 0], [0, 0], [0, 0.15]], np.float32))
    B, seed = 6000, 77
    t, out = simulate.run_point_dl(handle, 2.5, B, tep_info, taps, 0.05, net.W1, net.W2, seed=seed, chunk=4096)
    # module chain on the same Philox frames
    yd = empty((B, 128), torch.float32)
    td = empty((B, 4), torch.int32)
    handle.call("ldpcb_gen_frames", seed, 0, B, 2.5, yd, td, None)
    sync()
    y = yd.cpu().numpy()
    lab = _lib.unpack_bits(td.cpu().numpy().view(np.uint32)).astype(np.int64)
    model = ms_test.Decoding_model()
    fer, ber, und, buffer = model(y, lab)
    assert t.frames == B and t.nms_frame_err == round(fer * B) and t.nms_undetected == und
    rows = np.stack(buffer[0])
    labs = np.stack(buffer[1])
    nn = nn_net.conv_bitwise()
    nn.set_taps(taps, 0.05)
    squashed, inputs0, labels0 = nn.preprocessing_inputs((rows, labs))
    s, f, w, c = osd.sliding_osd(net, rows, nn(squashed), labels0, tep_info)
    assert (out["dl_success"], out["dl_failure"], out["windows_sum"], out["complexity_sum"]) == (s, f, w, c)
    assert s + f == t.nms_detected and out["fer_final"] == (f + und) / B


def test_every_kernel_under_the_bounds_check_build():
    """compute-sanitizer is closed on the GPU pool.  Instead: the library built with -DLDPCB_BOUNDS (device-side asserts on
    every data-dependent index: failure-list appends, inverse TEP tables, decoded candidate positions) runs
    scripts/sanitize_case.py -- every kernel at ragged sizes, quantised inputs for the exact fallbacks -- in a child
    process; a violated assert traps the kernel and the child exits non-zero."""
    import os
    import subprocess
    import sys

    from short_ldpc_decoding_osd_b200 import build

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = build.build(bounds=True)
    env = dict(os.environ, LDPCB_B200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "sanitize_case.py")], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "sanitize case done" in r.stdout
