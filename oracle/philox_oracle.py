"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the counter-based BPSK/AWGN frame generator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product (short_ldpc_decoding_osd_b200/) never does.

Restates (a) Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as
easy as 1, 2, 3", SC'11; Random123 v1.x philox.h -- an algorithm that is NOT in /root/reference; it is
pinned by the Random123 known-answer vectors in tests/test_oracle_framegen.py) and (b) the channel
model of the reference generator LDPC_128/Testing_data_gen_128/data_generating.py:13-51 (AWGN branch,
random codewords) on top of it, with the counter layout documented in include/ldpc_b200.h.
The reference itself draws from NumPy's MT19937 global stream, so there is no bit-level parity with
it; parity with the reference generator is statistical (tests compare moments and the FER it induces).
Normals are computed here in float64 and rounded to fp32; the kernel computes them in fp32
(logf, sqrtf, sincospif), so y agrees to a few ulp (tolerance 4e-6 abs in the tests), while the
message bits and the codeword are bit-exact.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [...,4] uint32, key [...,2] uint32 (broadcastable) -> [...,4] uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for _ in range(10):
        p0 = (np.uint64(M0) * c[0])
        p1 = (np.uint64(M1) * c[2])
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c = [(hi1 ^ c[1] ^ k0) & np.uint64(MASK), lo1, (hi0 ^ c[3] ^ k1) & np.uint64(MASK), lo0]
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return np.stack(c, axis=-1).astype(np.uint32)


def sigma_of(ebn0_db: float, k: int = 64, n: int = 128) -> np.float32:
    """data_generating.py:17, evaluated in float64 and rounded to fp32 as the library does."""
    return np.float32(np.sqrt(1.0 / (2.0 * (float(k) / float(n)) * 10.0 ** (float(np.float32(ebn0_db)) / 10.0))))


def _u01(x: np.ndarray) -> np.ndarray:
    """23 random bits + 1/2 (the Box-Muller angle)."""
    return ((x >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0 ** -23


def _u01_32(x: np.ndarray) -> np.ndarray:
    """The Box-Muller radius uniform from all 32 bits, defined in fp32 exactly as csrc/philox.cuh u01_32 computes it:
    fl32(fl32(float(x) + 0.5) * 2^-32), clamped to the largest float below 1 (tail out to 6.76 sigma)."""
    u = (x.astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -32)
    return np.minimum(u, np.float32(0.99999994)).astype(np.float64)


def gen_frames(seed: int, first_frame: int, B: int, ebn0_db: float, G: np.ndarray):
    """-> (y float32 [B,128], codewords uint8 [B,128], message bits uint8 [B,64])."""
    G = np.asarray(G).astype(np.int64) & 1
    k, n = G.shape
    f = (np.arange(B, dtype=np.uint64) + np.uint64(first_frame))
    flo = (f & np.uint64(MASK)).astype(np.uint32)
    fhi = (f >> np.uint64(32)).astype(np.uint32)
    key = np.array([seed & MASK, (seed >> 32) & MASK], dtype=np.uint32)
    # message bits: stream 1, block 0
    ctr = np.stack([flo, fhi, np.zeros(B, np.uint32), np.ones(B, np.uint32)], axis=-1)
    mw = philox4x32_10(ctr, key)
    msg64 = mw[:, 0].astype(np.uint64) | (mw[:, 1].astype(np.uint64) << np.uint64(32))
    msg = ((msg64[:, None] >> np.arange(64, dtype=np.uint64)[None]) & np.uint64(1)).astype(np.int64)
    cw = (msg.dot(G) % 2).astype(np.uint8)
    # normals: stream 0, blocks 0..31
    blk = np.arange(32, dtype=np.uint32)
    ctr = np.stack([np.repeat(flo, 32), np.repeat(fhi, 32), np.tile(blk, B), np.zeros(B * 32, np.uint32)], axis=-1)
    x = philox4x32_10(ctr, key)  # [B*32,4]
    u, ur = _u01(x), _u01_32(x)
    z = np.empty((B * 32, 4), dtype=np.float64)
    for a, b in ((0, 1), (2, 3)):
        r = np.sqrt(-2.0 * np.log(ur[:, a]))
        z[:, a] = r * np.cos(2.0 * np.pi * u[:, b])
        z[:, b] = r * np.sin(2.0 * np.pi * u[:, b])
    z = z.reshape(B, n)
    sigma = float(sigma_of(ebn0_db, k, n))
    ch = 1.0 + sigma * z
    y = np.where(cw == 0, ch, -ch).astype(np.float32)
    return y, cw, msg.astype(np.uint8)
