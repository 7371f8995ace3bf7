"""Drop-in for ``LDPC_128/Testing_data_gen_128/data_generating.py``: BPSK/AWGN test frames.

``testing_data_generating(code, SNR, max_frame, seed=0, first_frame=0)`` -> ``(testing_data float32[F,128],
testing_data_labels int64[F,128])`` with the reference's channel model (``:13-51``, AWGN branch, random
codewords, no LLR scaling).  The reference draws from NumPy's unseeded global MT19937 stream; here the frames
come from the counter-based Philox kernel (ldpcb_gen_frames), so frame f of a run depends only on (seed, f) and
any shard can be generated on any GPU.  Rayleigh fading (``:21-38``) is out of scope (off in every driver).

``replay_numpy_seed=s`` replays the reference's own noise instead: the frames ``np.random.seed(s)`` followed by the
reference's ``testing_data_generating`` would produce (legacy MT19937 stream: ``normal(1, sigma, [F, n])`` drawn
first, then ``randint(0, 2, [F, k])``, ``:40-45``), bit for bit -- the way to hand both decoders identical inputs
in a parity run.  A sequential MT19937 stream cannot be sharded, so this mode is host NumPy by nature (it is the
reference's generator, not a decoder path) and is not what the Monte-Carlo harness uses.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from . import globalmap as GL
from .runtime import get_handle


def testing_data_generating(code, SNR, max_frame, seed: int = 0, first_frame: int = 0, replay_numpy_seed=None):
    if GL.map.get("Rayleigh_fading"):
        raise NotImplementedError("Rayleigh fading is not part of the hot path (Main_test.py:34 sets it False)")
    if replay_numpy_seed is not None:
        n, k = int(code.check_matrix_column), int(code.k)
        F = int(max_frame)
        rs = np.random.RandomState(int(replay_numpy_seed))  # the stream np.random.seed(s) gives the reference
        sigma = np.sqrt(1.0 / (2 * (float(k) / float(n)) * 10 ** (SNR / 10)))
        channel = rs.normal(1, sigma, size=(F, n))           # data_generating.py:40 -- drawn before the messages
        if GL.map.get("ALL_ZEROS_CODEWORD_TESTING"):
            return channel, np.zeros((F, n), dtype=np.int64)
        msg = rs.randint(0, 2, size=[F, k], dtype=int)       # :42
        cw = msg.dot(np.asarray(code.G)) % 2                 # :43
        return np.where(cw == 0, channel, -channel), cw      # :44-45 (float64, like the reference; its files store float32)
    import torch  # device memory carrier only

    h = get_handle(code)
    dev = f"cuda:{h.device}"
    F = int(max_frame)
    y = torch.empty((F, 128), dtype=torch.float32, device=dev)
    cw = torch.empty((F, 4), dtype=torch.int32, device=dev)
    h.call("ldpcb_gen_frames", int(seed), int(first_frame), F, float(SNR), y, cw, None)
    torch.cuda.synchronize()
    labels = _lib.unpack_bits(cw.cpu().numpy().view(np.uint32)).astype(np.int64)
    data = y.cpu().numpy()
    if GL.map.get("ALL_ZEROS_CODEWORD_TESTING"):
        data = np.where(labels == 0, data, -data)  # same noise, all-zero codeword
        labels = np.zeros_like(labels)
    return data, labels
