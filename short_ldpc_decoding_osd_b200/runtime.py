"""Process-wide decoder handle for the drop-in modules (one ldpcb handle per (process, device))."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib
from . import globalmap as GL

_handles: Dict[Tuple[int, bytes], _lib.Handle] = {}


def current_device() -> int:
    return int(os.environ.get("LDPCB_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def get_handle(code=None, device: Optional[int] = None) -> _lib.Handle:
    """The handle for `code` (default: GL.get_map('code_parameters')) on `device`; created on first use.
    Raises if the CUDA library cannot be loaded or no GPU exists -- there is no CPU fallback."""
    if code is None:
        code = GL.map.get("code_parameters")
        if code is None:
            from .fill_matrix_info import Code

            code = Code()
            GL.set_map("code_parameters", code)
    dev = current_device() if device is None else device
    key = (dev, np.asarray(code.H, dtype=np.uint8).tobytes())
    h = _handles.get(key)
    if h is None:
        h = _lib.Handle(code.H, code.G, device=dev)
        _handles[key] = h
    return h


def close_all() -> None:
    for h in _handles.values():
        h.close()
    _handles.clear()


def softplus(x) -> float:
    """tf.nn.softplus evaluated in fp32 (ms_test.py:207-208)."""
    x = np.float32(x)
    return float(np.log1p(np.exp(x, dtype=np.float32), dtype=np.float32))
