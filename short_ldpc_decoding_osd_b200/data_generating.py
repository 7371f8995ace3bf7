"""Drop-in for ``LDPC_128/Testing_data_gen_128/data_generating.py``: BPSK/AWGN test frames.

``testing_data_generating(code, SNR, max_frame, seed=0, first_frame=0)`` -> ``(testing_data float32[F,128],
testing_data_labels int64[F,128])`` with the reference's channel model (``:13-51``, AWGN branch, random
codewords, no LLR scaling).  The reference draws from NumPy's unseeded global MT19937 stream; here the frames
come from the counter-based Philox kernel (ldpcb_gen_frames), so frame f of a run depends only on (seed, f) and
any shard can be generated on any GPU.  Rayleigh fading (``:21-38``) is out of scope (off in every driver).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from . import globalmap as GL
from .runtime import get_handle


def testing_data_generating(code, SNR, max_frame, seed: int = 0, first_frame: int = 0):
    if GL.map.get("Rayleigh_fading"):
        raise NotImplementedError("Rayleigh fading is not part of the hot path (Main_test.py:34 sets it False)")
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
