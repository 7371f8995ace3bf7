"""-m gpu: the quasi-cyclic NMS kernel (csrc/nms_qc.cu, half a warp per frame, shuffle exchanges) against the
table-driven kernel (csrc/nms.cu) and the oracle, and the fused device pipelines against their unfused composition.
Everything through the C ABI; float outputs are compared BIT FOR BIT (both kernels restate the reference's fp32
operations in the same order: ms_test.py:124-137,180-228)."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import philox_oracle as PO
from short_ldpc_decoding_osd_b200 import _lib
from tests.gpu_util import dev, empty, nms_gpu, sync

pytestmark = pytest.mark.gpu
ALPHA = 0.66943514


@pytest.fixture(scope="module")
def generic_handle(code):
    """A handle whose NMS runs on the table-driven shared-memory kernel and whose pipelines are unfused."""
    os.environ["LDPCB_NMS_GENERIC"] = "1"
    try:
        h = _lib.Handle(code.H, code.G, device=0)
    finally:
        del os.environ["LDPCB_NMS_GENERIC"]
    yield h
    h.close()


def _same_floats(a, b):
    """Exact equality of fp32 values; the only bit patterns allowed to differ are the signs of zeros (the kernels carry
    the sign product through a zeroed check, the oracle multiplies by tf.sign(0) = 0: both are zero messages)."""
    ua, ub = np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)
    diff = ua != ub
    return bool(np.all(a[diff] == 0.0) and np.all(b[diff] == 0.0))


def _edge_frames(code, B, seed):
    y, cw, _ = PO.gen_frames(seed, 0, B, 2.5, code.G)
    y = y.copy()
    y[0] = 0.0                      # tf.sign(0) = 0 zeroes whole checks
    y[1, :9] = 0.0
    y[2] *= 1e20                    # clip at 1e30 after a few iterations
    y[3] *= 1e-30                   # denormal messages
    y[4] = -0.0
    y[5] = np.where(cw[5] == 0, 1.0, -1.0)
    y[6, ::3] = np.float32(1e-45)
    return y, cw


@pytest.mark.parametrize("B", [1, 2, 3, 33, 1187, 70001])
def test_qc_kernel_equals_generic_kernel_bit_for_bit(handle, generic_handle, code, B):
    y, _ = _edge_frames(code, max(B, 8), 11)
    y = y[:B]
    traj = B <= 1187
    a = nms_gpu(handle, y, traj=traj)
    b = nms_gpu(generic_handle, y, traj=traj)
    assert np.array_equal(a["hard"], b["hard"])
    assert np.array_equal(a["syndrome_nz"], b["syndrome_nz"])
    assert np.array_equal(a["iters_used"], b["iters_used"])
    if traj:
        assert np.array_equal(a["traj"].view(np.uint32), b["traj"].view(np.uint32)), "posteriors differ in some bit"


def test_qc_kernel_equals_oracle(handle, code):
    y, _ = _edge_frames(code, 3000, 5)
    g = nms_gpu(handle, y, traj=True)
    ref = CO.nms(y, code.H, 12, ALPHA, traj=True)
    assert np.array_equal(g["hard"], ref["hard"])
    assert np.array_equal(g["syndrome_nz"], ref["syndrome_nz"])
    assert _same_floats(g["traj"], ref["traj"])


def test_qc_kernel_nms2_weight_and_iteration_counts(handle, generic_handle, code):
    y, _ = _edge_frames(code, 500, 9)
    for iters, w in ((1, 1.0), (5, 0.83), (20, 1.25)):
        a = nms_gpu(handle, y, iters=iters, alpha=0.7, w_vc=w, w_marg=w)
        b = nms_gpu(generic_handle, y, iters=iters, alpha=0.7, w_vc=w, w_marg=w)
        assert np.array_equal(a["hard"], b["hard"]) and np.array_equal(a["traj"].view(np.uint32), b["traj"].view(np.uint32))
        ref = CO.nms(y, code.H, iters, 0.7, w, w, traj=True)
        assert _same_floats(a["traj"], ref["traj"])


def test_qc_fir_variant_equals_generic(handle, generic_handle, code):
    y, _ = _edge_frames(code, 777, 3)
    taps = np.linspace(-0.3, 0.4, 13).astype(np.float32)
    out = []
    for h in (handle, generic_handle):
        yd = dev(y)
        bits, syn, met = empty((777, 4), torch.int32), empty((777,), torch.uint8), empty((777, 128), torch.float32)
        h.call("ldpcb_nms_decode_fir", yd, 777, 12, ALPHA, 1.0, 1.0, taps, 0.05, bits, syn, met, None)
        sync()
        out.append((bits.cpu().numpy(), syn.cpu().numpy(), met.cpu().numpy()))
    for u, v in zip(out[0], out[1]):
        assert np.array_equal(u.view(np.uint8), v.view(np.uint8))


@pytest.mark.parametrize("order,B", [(2, 20001), (1, 4099), (0, 513), (3, 300), (-1, 1000), (2, 1), (2, 2), (3, 3), (1, 7)])
def test_fused_decode_equals_unfused_pipeline(handle, generic_handle, code, order, B):
    """ldpcb_decode on the fused path (NMS + tallies + failure list in one kernel, OSD + tallies in the other) gives
    the same decisions, flags, TEP choices and all 16 counters as the 7-launch pipeline."""
    y, cw = _edge_frames(code, max(B, 8), 21)
    y, cw = np.ascontiguousarray(y[:B]), np.ascontiguousarray(cw[:B])
    truth = dev(_lib.pack_bits(cw).view(np.int32))
    res = []
    for h in (handle, generic_handle):
        yd = dev(y)
        bits, syn, bt = empty((B, 4), torch.int32), empty((B,), torch.uint8), empty((B,), torch.int32)
        cnt = torch.zeros(16, dtype=torch.int64, device="cuda:0")
        for _ in range(2):  # accumulation
            h.call("ldpcb_decode", yd, B, 12, ALPHA, 1.0, 1.0, 0, order, 0, bits, syn, bt, truth, cnt, None)
        sync()
        res.append((bits.cpu().numpy(), syn.cpu().numpy(), bt.cpu().numpy(), cnt.cpu().numpy()))
    for name, u, v in zip(("bits", "syndrome", "best_tep", "counters"), res[0], res[1]):
        assert np.array_equal(u, v), name
    assert res[0][3][0] == 2 * B


def test_fused_decode_without_truth_or_best_tep(handle, generic_handle, code):
    y, _ = _edge_frames(code, 2500, 2)
    out = []
    for h in (handle, generic_handle):
        yd = dev(y)
        bits = empty((2500, 4), torch.int32)
        h.call("ldpcb_decode", yd, 2500, 12, ALPHA, 1.0, 1.0, 0, 2, 1, bits, None, None, None, None, None)
        sync()
        out.append(bits.cpu().numpy())
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("order,ebn0,B", [(2, 2.5, 30001), (1, 1.5, 5000), (-1, 3.0, 999)])
def test_fused_simulate_equals_unfused(handle, generic_handle, order, ebn0, B):
    """The Monte-Carlo step with the generator fused into the decoder's prologue tallies exactly what
    generate -> decode -> tally does (same Philox frames, same decisions)."""
    cs = []
    for h in (handle, generic_handle):
        cnt = torch.zeros(16, dtype=torch.int64, device="cuda:0")
        h.call("ldpcb_simulate", 77, 1 << 40, B, ebn0, 12, ALPHA, 1.0, 1.0, 0, order, 0, cnt, None)
        sync()
        cs.append(cnt.cpu().numpy())
    assert np.array_equal(cs[0], cs[1]), (cs[0], cs[1])
    assert cs[0][0] == B


def test_two_streams_do_not_share_scratch(handle, code):
    """ldpcb_decode calls of one handle on two streams used to carve the same workspace."""
    B = 40000
    y1, c1 = _edge_frames(code, B, 31)
    y2, c2 = _edge_frames(code, B, 32)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    bufs = []
    for y, cw, st in ((y1, c1, s1), (y2, c2, s2)):
        yd, tr = dev(y), dev(_lib.pack_bits(cw).view(np.int32))
        bits = empty((B, 4), torch.int32)
        cnt = torch.zeros(16, dtype=torch.int64, device="cuda:0")
        bufs.append((yd, tr, bits, cnt, st))
    sync()
    for rep in range(3):
        for yd, tr, bits, cnt, st in bufs:
            handle.call("ldpcb_decode", yd, B, 12, ALPHA, 1.0, 1.0, 0, 2, 0, bits, None, None, tr, cnt, st.cuda_stream)
    sync()
    for yd, tr, bits, cnt, st in bufs:
        outs.append((bits.cpu().numpy().copy(), cnt.cpu().numpy().copy()))
    # reference: the same calls one after the other on the default stream
    for (yd, tr, bits, cnt, st), (b_par, c_par) in zip(bufs, outs):
        cnt.zero_()
        for rep in range(3):
            handle.call("ldpcb_decode", yd, B, 12, ALPHA, 1.0, 1.0, 0, 2, 0, bits, None, None, tr, cnt, None)
        sync()
        assert np.array_equal(bits.cpu().numpy(), b_par) and np.array_equal(cnt.cpu().numpy(), c_par)


@pytest.mark.parametrize("B,cap", [(1000, 1000), (20000, 25000), (3000, 40)])
def test_retest_host_equals_decode_then_trajectories(handle, generic_handle, code, B, cap):
    """ldpcb_nms_retest_host (Decoding_model.call in one call): tallies, flags and the 13-row records of the detected
    failures equal a full-trajectory decode of the batch; a too-small record buffer reports the true count."""
    y, cw = _edge_frames(code, B, 41)
    truth = _lib.pack_bits(cw)
    full = nms_gpu(handle, y, traj=True)
    want_idx = np.flatnonzero(full["syndrome_nz"])
    for h in (handle, generic_handle):
        bits, syn = np.empty((B, 4), np.uint32), np.empty(B, np.uint8)
        cnt = np.zeros(16, np.uint64)
        fidx, traj, nf = np.empty(cap, np.int32), np.empty((cap, 13, 128), np.float32), np.zeros(1, np.int64)
        h.call("ldpcb_nms_retest_host", y, B, 12, ALPHA, 1.0, 1.0, truth, bits, syn, cnt, cap, fidx, traj, nf)
        n = int(nf[0])
        assert n == len(want_idx)
        k = min(n, cap)
        assert np.array_equal(fidx[:k], want_idx[:k])
        assert np.array_equal(traj[:k].view(np.uint32), full["traj"][want_idx[:k]].view(np.uint32))
        assert np.array_equal(_lib.unpack_bits(bits), full["hard"]) and np.array_equal(syn.astype(bool), full["syndrome_nz"])
        wrong = (full["hard"] != cw).any(1)
        assert int(cnt[0]) == B and int(cnt[1]) == int(wrong.sum()) and int(cnt[2]) == int((full["hard"] != cw).sum())
        assert int(cnt[3]) == n and int(cnt[4]) == int((wrong & ~full["syndrome_nz"]).sum()) and int(cnt[9]) == int(cnt[1])


def test_handles_of_two_devices_in_one_thread(code):
    """every entry point makes its handle's device current (and restores the caller's), so handles of several
    devices can be interleaved from one thread.  Needs two GPUs; skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    B = 5000
    y, cw = _edge_frames(code, B, 51)
    truth = _lib.pack_bits(cw).view(np.int32)
    hs = [_lib.Handle(code.H, code.G, device=d) for d in (0, 1)]
    assert torch.cuda.current_device() == 0  # ldpcb_create restored the caller's device
    bufs = []
    for d in (0, 1):
        dev_ = f"cuda:{d}"
        bufs.append((torch.from_numpy(y).to(dev_), torch.from_numpy(truth).to(dev_), torch.empty((B, 4), dtype=torch.int32, device=dev_),
                     torch.zeros(16, dtype=torch.int64, device=dev_)))
    for rep in range(3):  # alternate devices without ever calling cudaSetDevice ourselves
        for h, (yd, tr, bits, cnt) in zip(hs, bufs):
            h.call("ldpcb_decode", yd, B, 12, ALPHA, 1.0, 1.0, 0, 2, 0, bits, None, None, tr, cnt, None)
            assert torch.cuda.current_device() == 0
    for d in (0, 1):
        torch.cuda.synchronize(d)
    r0, r1 = bufs[0][2].cpu().numpy(), bufs[1][2].cpu().numpy()
    assert np.array_equal(r0, r1) and np.array_equal(bufs[0][3].cpu().numpy(), bufs[1][3].cpu().numpy())
    assert int(bufs[0][3][0].item()) == 3 * B
    for h in hs:
        h.close()
