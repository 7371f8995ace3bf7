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
OSD_MAXW_SHIFT = 4

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
    "ldpcb_nms_retest_host": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "ldpcb_decode_host": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _f32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ldpcb_device_pci_bus_id": (_i32, [_i32, C.c_char_p, _i32]),
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


class _PinnedBlock:
    """Owner of one cudaMallocHost block.  Every array handed out is a view whose ultimate ``.base`` is the
    ndarray built from this object's ``__array_interface__``, and that ndarray keeps this object alive -- so the
    block is freed only when the last view (slice, reshape, .view()) of it is gone."""

    def __init__(self, lib, nbytes: int):
        p = _vp()
        st = lib.ldpcb_host_alloc(C.byref(p), max(nbytes, 1))
        if st != 0:
            raise LdpcB200Error(st, lib.ldpcb_last_error(None).decode())
        self._lib, self.ptr, self.nbytes = lib, p.value, nbytes
        self.__array_interface__ = {"data": (self.ptr, False), "shape": (max(nbytes, 1),), "typestr": "|u1", "version": 3}

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                self._lib.ldpcb_host_free(ptr)
            except Exception:
                pass


class _PinnedPool:
    """Recycles pinned blocks by size class (powers of two from 64 KiB): cudaMallocHost costs ~0.3 ms per call, more than a
    1000-frame decode, so the drop-in entry points that return fresh arrays per call take their result buffers from here.
    A block goes back to the pool when the last array viewing it is collected; the pool keeps at most `limit` bytes."""

    def __init__(self, limit: int = 1 << 30):
        self.free, self.held, self.limit = {}, 0, limit

    def take(self, nbytes: int):
        size = 1 << max(16, (max(nbytes, 1) - 1).bit_length())
        lst = self.free.get(size)
        if lst:
            self.held -= size
            return _PooledBlock(self, lst.pop(), size)
        return _PooledBlock(self, _PinnedBlock(load(), size), size)

    def give(self, raw, size: int):
        if self.held + size <= self.limit:
            self.free.setdefault(size, []).append(raw)
            self.held += size


class _PooledBlock:
    def __init__(self, pool, raw, size):
        self._pool, self._raw, self._size = pool, raw, size
        self.__array_interface__ = raw.__array_interface__

    def __del__(self):
        try:
            self._pool.give(self._raw, self._size)
        except Exception:
            pass


_POOL = _PinnedPool()


def pinned_pool_empty(shape, dtype) -> np.ndarray:
    """Like pinned_empty, from the recycling pool."""
    return pinned_empty(shape, dtype, _alloc=_POOL.take)


def pinned_empty(shape, dtype, _alloc=None) -> np.ndarray:
    """NumPy array backed by pinned host memory (ldpcb_host_alloc); the memory is released when the array AND
    every view derived from it have been collected.  `_alloc` (tests) substitutes the block allocator."""
    dt = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = count * dt.itemsize
    block = (_alloc or (lambda n: _PinnedBlock(load(), n)))(nbytes)
    raw = np.asarray(block)  # base -> block
    return raw[:nbytes].view(dt).reshape(shape)


def device_numa_node(device: int) -> int:
    """NUMA node of a CUDA device from sysfs, -1 if unknown (single-node hosts report -1 or 0)."""
    buf = C.create_string_buffer(32)
    if load().ldpcb_device_pci_bus_id(int(device), buf, 32) != 0:
        return -1
    try:
        with open(f"/sys/bus/pci/devices/{buf.value.decode().lower()}/numa_node") as f:
            return int(f.read().strip())
    except (OSError, ValueError):
        return -1


def bind_host_to_device(device: int) -> dict:
    """Restrict this process to the CPUs of the NUMA node `device` hangs off, so that pinned buffers allocated (and
    first touched) afterwards live in the memory closest to the GPU's PCIe root port.  A no-op on hosts with one
    node.  Returns what was done (bench.py prints it)."""
    node = device_numa_node(device)
    info = {"numa_node": node, "bound": False}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        if node < 0 or len(nodes) <= 1:
            return info
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
    except (OSError, ValueError, AttributeError) as e:
        info["error"] = str(e)[:100]
    return info


def unpack_bits(words: np.ndarray) -> np.ndarray:
    """[B,4] uint32 little-endian packed frames -> [B,128] uint8 bits."""
    w = np.ascontiguousarray(words, dtype="<u4")
    return np.unpackbits(w.view(np.uint8).reshape(w.shape[0], 16), axis=1, bitorder="little")


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """[B,128] 0/1 -> [B,4] uint32 little-endian packed frames."""
    b = np.ascontiguousarray(np.asarray(bits) & 1, dtype=np.uint8)
    return np.packbits(b, axis=1, bitorder="little").view("<u4").reshape(b.shape[0], 4)
