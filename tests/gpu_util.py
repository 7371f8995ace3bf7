"""Helpers shared by the -m gpu parity tests: torch is only the device-memory carrier."""
import numpy as np
import torch

from short_ldpc_decoding_osd_b200 import _lib

DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=DEV)


def sync():
    torch.cuda.synchronize()


def nms_gpu(h, y, iters=12, alpha=0.66943514, w_vc=1.0, w_marg=1.0, early=0, traj=True):
    B = y.shape[0]
    yd = dev(y.astype(np.float32))
    bits = empty((B, 4), torch.int32)
    it = empty((B,), torch.uint8)
    syn = empty((B,), torch.uint8)
    tr = empty((B, iters + 1, 128), torch.float32) if traj else None
    h.call("ldpcb_nms_decode", yd, B, iters, float(alpha), float(w_vc), float(w_marg), int(early), bits, it, syn, tr, None)
    sync()
    return {
        "hard": _lib.unpack_bits(bits.cpu().numpy().view(np.uint32)),
        "iters_used": it.cpu().numpy(),
        "syndrome_nz": syn.cpu().numpy().astype(bool),
        "traj": tr.cpu().numpy() if traj else None,
    }


def osd_gpu(h, yo, ys=None, order=1, tep_order=0, flags=0):
    B = yo.shape[0]
    yod = dev(yo.astype(np.float32))
    ysd = yod if ys is None else dev(ys.astype(np.float32))
    bits = empty((B, 4), torch.int32)
    bt = empty((B,), torch.int32)
    bq = empty((B,), torch.int64)
    ex = empty((B,), torch.int32)
    pm = empty((B, 128), torch.uint8)
    rg = empty((B, 64), torch.int64)
    h.call("ldpcb_osd_decode", yod, ysd, B, order, tep_order, flags, bits, bt, bq, ex, pm, rg, None)
    sync()
    return {
        "codeword": _lib.unpack_bits(bits.cpu().numpy().view(np.uint32)),
        "best_tep": bt.cpu().numpy(),
        "best_score_q": bq.cpu().numpy(),
        "score_exp": ex.cpu().numpy(),
        "perm": pm.cpu().numpy(),
        "redG": rg.cpu().numpy().view(np.uint64),
    }


def redG_to_matrix(redg_row_words):
    """[64] uint64 P' words -> int[64,128] reduced_G = [I | P']."""
    P = ((redg_row_words[:, None] >> np.arange(64, dtype=np.uint64)[None]) & np.uint64(1)).astype(np.int64)
    return np.concatenate([np.identity(64, dtype=np.int64), P], axis=1)
