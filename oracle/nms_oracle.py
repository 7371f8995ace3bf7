"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the normalized min-sum decoder.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product (short_ldpc_decoding_osd_b200/) never does.

NumPy restatement, op by op, of the reference's dense TensorFlow graph in
LDPC_128/Ldpc_128_testing/ms_test.py.  Parity status: the reference ships no tests or golden
vectors for this path; this restatement is pinned by
tests/golden/nms_ref_shim.npz, produced by running the reference's own ms_test.py source under a
NumPy emulation of the TensorFlow ops it calls (oracle/ref_runner.py) -- TensorFlow itself is not
installable here.

TF semantics restated (the ones that matter):
  tf.sign(0) = 0                       -> a zero variable-to-check message zeroes its whole check
  tf.nn.top_k(x, 2) values             -> two largest, duplicates kept (min2 == min1 on a tie)
  tf.where(a > b, x, y)                -> strict comparison
  tf.reduce_sum(cv, axis=1) fp32       -> order unspecified in TF; defined here as ascending check
                                          index, sequentially (adding the zeros of absent edges is
                                          exact, so this equals summing the live edges in that order)
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def softplus(x: float) -> np.float32:
    """tf.nn.softplus in fp32 (ms_test.py:207-208); softplus(-0.048) = 0.66943514."""
    x = F32(x)
    return F32(np.log1p(np.exp(x, dtype=F32), dtype=F32))


def compute_vc(cv, y, Hf, w_vc):
    # ms_test.py:124-137
    soft_input_weighted = y * F32(w_vc)
    temp = np.zeros_like(y)
    for c in range(cv.shape[1]):  # tf.reduce_sum(cv_matrix, 1), ascending check order
        temp = temp + cv[:, c, :]
    temp = temp + soft_input_weighted
    temp = temp[:, None, :] * Hf[None]
    return temp - cv


def compute_cv2(vc, H, Hf, alpha):
    # ms_test.py:180-210
    supplement = (1 - H).astype(F32)[None]
    vc_sign = np.sign(supplement + vc).astype(F32)
    temp1 = np.prod(vc_sign, axis=2, dtype=F32)[:, :, None]
    result_sign = (temp1 * Hf[None]) * vc_sign
    back = np.where(H == 0, F32(-1e30 - 1), F32(0.0)).astype(F32)[None]
    vc_abs_clip = np.clip(np.abs(vc), F32(0), F32(1e30))
    decision = -np.abs(vc_abs_clip) + back
    top2 = -np.sort(-decision, axis=2)[:, :, :2]  # tf.nn.top_k(k=2): two largest, descending
    min1 = (-top2[:, :, 0])[:, :, None] * Hf[None]
    min2 = (-top2[:, :, 1])[:, :, None] * Hf[None]
    result = np.where(vc_abs_clip > min1, min1, min2)
    return (F32(alpha) * result) * result_sign


def marginalize(cv, y, w_marg):
    # ms_test.py:220-228
    temp = np.zeros_like(y)
    for c in range(cv.shape[1]):
        temp = temp + cv[:, c, :]
    return temp + F32(w_marg) * y


def belief_propagation(y, H, iters=12, alpha=None, w_vc=1.0, w_marg=1.0):
    """ms_test.py:106-121,234-242 -> list of iters+1 soft outputs (index 0 = input)."""
    y = np.ascontiguousarray(y, dtype=F32)
    if alpha is None:
        alpha = softplus(-0.048)
    H = np.asarray(H).astype(np.int64)
    Hf = H.astype(F32)
    B = y.shape[0]
    cv = np.zeros((B, H.shape[0], H.shape[1]), dtype=F32)
    out = [y]
    for _ in range(iters):
        vc = compute_vc(cv, y, Hf, w_vc)
        cv = compute_cv2(vc, H, Hf, alpha)
        out.append(marginalize(cv, y, w_marg))
    return out


def hard_decision(soft):
    """tf.where(soft > 0, 0, 1) (ms_test.py:39): 0.0 and NaN map to bit 1."""
    return np.where(soft > 0, 0, 1).astype(np.int64)


def get_eval(soft_output_list, labels, H):
    """ms_test.py:36-54 -> (FER, BER, n_undetected, index[F,1], hard[B,128], syndrome_nz[B])."""
    hard = hard_decision(soft_output_list[-1])
    labels = np.asarray(labels).astype(np.int64)
    err_bit_sum = (hard != labels).sum(axis=-1)
    syndrome = (hard.dot(np.asarray(H).T) % 2).sum(axis=-1)
    success = err_bit_sum == 0
    undetected = int(np.sum((syndrome == 0) & ~success))
    index = np.flatnonzero(syndrome != 0)[:, None]
    fer = 1 - success.sum() / labels.shape[0]
    ber = err_bit_sum.sum() / (labels.shape[0] * labels.shape[1])
    return fer, ber, undetected, index, hard, (syndrome != 0)


def decode(y, H, iters=12, alpha=None, w_vc=1.0, w_marg=1.0, early_stop=False, chunk=500):
    """Batched driver -> dict(hard, syndrome_nz, iters_used, traj[B,iters+1,128]).

    early_stop mirrors the kernel's optional mode (not in the reference): a frame freezes at the
    first iteration whose hard decision has a zero syndrome; later trajectory rows repeat it.
    """
    y = np.ascontiguousarray(y, dtype=F32)
    Hm = np.asarray(H)
    B = y.shape[0]
    traj = np.empty((B, iters + 1, y.shape[1]), dtype=F32)
    for s in range(0, B, chunk):
        outs = belief_propagation(y[s:s + chunk], Hm, iters, alpha, w_vc, w_marg)
        traj[s:s + chunk] = np.stack(outs, axis=1)
    iters_used = np.full(B, iters, dtype=np.uint8)
    if early_stop:
        done = np.zeros(B, dtype=bool)
        for it in range(1, iters + 1):
            hard = hard_decision(traj[:, it])
            ok = (hard.dot(Hm.T) % 2).sum(axis=-1) == 0
            newly = ok & ~done
            iters_used[newly] = it
            for later in range(it + 1, iters + 1):
                traj[newly, later] = traj[newly, it]
            done |= ok
    hard = hard_decision(traj[:, iters])
    syn = (hard.dot(Hm.T) % 2).sum(axis=-1) != 0
    return {"hard": hard.astype(np.uint8), "syndrome_nz": syn, "iters_used": iters_used, "traj": traj}
