"""CPU tests of the drop-in boundary: the shared library builds, loads and exports exactly the symbols
include/ldpc_b200.h declares (no compute calls without a GPU), and fails loudly where it must."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from short_ldpc_decoding_osd_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldpcb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load(build_if_missing=False)
    declared = header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ldpc_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes prototypes and header disagree"
    assert lib.ldpcb_abi_version() == 1


def test_library_is_sm100a_only():
    out = os.popen(f"cuobjdump -lelf {build.LIB_PATH} 2>/dev/null").read()
    if out.strip():
        assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_create_fails_loudly_without_device_or_with_bad_code(code):
    lib = _lib.load()
    H = np.ascontiguousarray(code.H, dtype=np.uint8)
    G = np.ascontiguousarray(code.G, dtype=np.uint8)
    h = C.c_void_p()
    if lib.ldpcb_device_count() == 0:
        st = lib.ldpcb_create(C.byref(h), H.ctypes.data, G.ctypes.data, 128, 64, 64, 0)
        assert st == -6 and b"no CPU fallback" in lib.ldpcb_last_error(None)
        with pytest.raises(_lib.LdpcB200Error):
            _lib.Handle(code.H, code.G)
    st = lib.ldpcb_create(C.byref(h), H.ctypes.data, G.ctypes.data, 96, 48, 48, 0)
    assert st == -2 and b"only n=128" in lib.ldpcb_last_error(None)
    assert lib.ldpcb_create(None, H.ctypes.data, G.ctypes.data, 128, 64, 64, 0) == -1
    assert lib.ldpcb_tep_count(None, 1, 0) == -1
    assert lib.ldpcb_launch_count(None) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "short_ldpc_decoding_osd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "libldpc_oracle" not in text and "c_oracle" not in text, f"{f} loads the C oracle"
                assert "/root/reference" not in text, f"{f} reads the reference at run time"


def test_pack_unpack_bits_roundtrip():
    rng = np.random.default_rng(0)
    b = rng.integers(0, 2, (9, 128)).astype(np.uint8)
    w = _lib.pack_bits(b)
    assert w.shape == (9, 4) and w.dtype == np.uint32
    assert np.array_equal(_lib.unpack_bits(w), b)
    j = 37
    assert ((int(w[0, j >> 5]) >> (j & 31)) & 1) == b[0, j]  # the header's bit convention
