"""ctypes binding of libldpc_b200.so (the C ABI declared in include/ldpc_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if it cannot be
loaded or no CUDA device exists the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build

N, M, K = 128, 64, 64
NUM_COUNTERS = 16
TEP_CONV, TEP_FS = 0, 1
OSD_TIES_HIGH_INDEX_FIRST = 1
OSD_DISC_HARD_FROM_SCORE = 2

COUNTER_NAMES = [
    "frames", "nms_frame_err", "nms_bit_err", "nms_detected", "nms_undetected", "nms_iters",
    "osd_frames", "osd_frame_err", "osd_bit_err", "final_frame_err", "final_bit_err", "teps",
    "phase0", "phase1", "phase2", "phase3",
]

# every symbol include/ldpc_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _u64, _f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
SYMBOLS = {
    "ldpcb_abi_version": (_i32, []),
    "ldpcb_device_count": (_i32, []),
    "ldpcb_create": (_i32, [C.POINTER(_vp), _vp, _vp, _i32, _i32, _i32, _i32]),
    "ldpcb_destroy": (None, [_vp]),
    "ldpcb_last_error": (C.c_char_p, [_vp]),
    "ldpcb_sm_count": (_i32, [_vp]),
    "ldpcb_gen_frames": (_i32, [_vp, _u64, _u64, _i64, _f32, _vp, _vp, _vp]),
    "ldpcb_nms_decode": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_osd_decode": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_osd_block_minima": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_tep_count": (_i32, [_vp, _i32, _i32]),
    "ldpcb_tep_table": (_i32, [_vp, _i32, _i32, _vp]),
    "ldpcb_select_flagged": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "ldpcb_gather_rows": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "ldpcb_dia_fir": (_i32, [_vp, _vp, _i64, _i32, _vp, _f32, _vp, _vp]),
    "ldpcb_nms_decode_fir": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _vp, _f32, _vp, _vp, _vp, _vp]),
    "ldpcb_dl_window_policy": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_tally": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "ldpcb_decode": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_simulate": (_i32, [_vp, _u64, _u64, _i64, _f32, _i32, _f32, _f32, _f32, _i32, _i32, _i32, _vp, _vp]),
    "ldpcb_nms_decode_host": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _i32, _vp, _vp, _vp, _vp]),
    "ldpcb_osd_decode_host": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_osd_sweep_host": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "ldpcb_osd_fs_decode": (_i32, [_vp, _vp, _i64, _i32, _f32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_osd_fs_decode_host": (_i32, [_vp, _vp, _i64, _i32, _f32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "ldpcb_osd_pb_decode": (_i32, [_vp, _vp, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_osd_pb_decode_host": (_i32, [_vp, _vp, _i64, _i32, _f32, _vp, _vp]),
    "ldpcb_decode_host": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_host_alloc": (_i32, [C.POINTER(_vp), _u64]),
    "ldpcb_host_free": (_i32, [_vp]),
    "ldpcb_launch_count": (_u64, [_vp]),
}

_LIB: Optional[C.CDLL] = None


class LdpcB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libldpc_b200 status {status}: {message}")
        self.status = status


def lib_path() -> str:
    """In-tree library, or the prebuilt one LDPCB_B200_LIB points at."""
    return os.environ.get("LDPCB_B200_LIB") or _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if necessary) the shared library and set the prototypes."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise FileNotFoundError(f"{path} is missing; run `python -m short_ldpc_decoding_osd_b200.build`")
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _ptr(x) -> Optional[int]:
    """Device pointer of a torch tensor / host pointer of a NumPy array / None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("NumPy arrays passed to libldpc_b200 must be C-contiguous")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensors passed to libldpc_b200 must be contiguous")
        return x.data_ptr()
    if isinstance(x, int):
        return x
    raise TypeError(f"cannot take a pointer of {type(x)}")


class Handle:
    """One decoder per (process, device): thin object wrapper over the C ABI."""

    def __init__(self, H: np.ndarray, G: np.ndarray, device: int = 0):
        self.lib = load()
        H8 = np.ascontiguousarray(np.asarray(H) & 1, dtype=np.uint8)
        G8 = np.ascontiguousarray(np.asarray(G) & 1, dtype=np.uint8)
        m, n = H8.shape
        k = G8.shape[0]
        self._h = _vp()
        st = self.lib.ldpcb_create(C.byref(self._h), H8.ctypes.data, G8.ctypes.data, n, m, k, device)
        if st != 0:
            msg = self.lib.ldpcb_last_error(None).decode()
            self._h = None
            raise LdpcB200Error(st, msg)
        self.device = device
        self.n, self.m, self.k = n, m, k

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.ldpcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int) -> None:
        if st != 0:
            raise LdpcB200Error(st, self.lib.ldpcb_last_error(self._h).decode())

    def call(self, name: str, *args) -> None:
        fn = getattr(self.lib, name)
        self._check(fn(self._h, *[(_ptr(a) if not isinstance(a, (int, float)) else a) for a in args]))

    # -- small conveniences --------------------------------------------------------------------
    @property
    def sm_count(self) -> int:
        return self.lib.ldpcb_sm_count(self._h)

    @property
    def launch_count(self) -> int:
        return int(self.lib.ldpcb_launch_count(self._h))

    def tep_count(self, order: int, tep_order: int = TEP_CONV) -> int:
        n = self.lib.ldpcb_tep_count(self._h, order, tep_order)
        if n < 0:
            self._check(n)
        return n

    def tep_table(self, order: int, tep_order: int = TEP_CONV) -> np.ndarray:
        out = np.empty(self.tep_count(order, tep_order), dtype=np.uint32)
        self._check(self.lib.ldpcb_tep_table(self._h, order, tep_order, out.ctypes.data))
        return out


def pinned_empty(shape, dtype) -> np.ndarray:
    """NumPy array backed by pinned host memory (ldpcb_host_alloc); freed when collected."""
    lib = load()
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize
    p = _vp()
    st = lib.ldpcb_host_alloc(C.byref(p), max(nbytes, 1))
    if st != 0:
        raise LdpcB200Error(st, lib.ldpcb_last_error(None).decode())
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    _PINNED[id(buf)] = (buf, p.value)
    import weakref

    weakref.finalize(arr, _free_pinned, id(buf))
    return arr


_PINNED = {}


def _free_pinned(key) -> None:
    ent = _PINNED.pop(key, None)
    if ent is not None and _LIB is not None:
        _LIB.ldpcb_host_free(ent[1])


def unpack_bits(words: np.ndarray) -> np.ndarray:
    """[B,4] uint32 little-endian packed frames -> [B,128] uint8 bits."""
    w = np.ascontiguousarray(words, dtype="<u4")
    return np.unpackbits(w.view(np.uint8).reshape(w.shape[0], 16), axis=1, bitorder="little")


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """[B,128] 0/1 -> [B,4] uint32 little-endian packed frames."""
    b = np.ascontiguousarray(np.asarray(bits) & 1, dtype=np.uint8)
    return np.packbits(b, axis=1, bitorder="little").view("<u4").reshape(b.shape[0], 4)
