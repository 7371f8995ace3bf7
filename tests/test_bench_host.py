"""CPU tests of the measurement harness and of the host-side helpers around the C ABI (no GPU)."""
import gc
import json
import os
import subprocess
import sys
import weakref

import numpy as np
import pytest

from short_ldpc_decoding_osd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _run_bench(args, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_uses_every_core_under_torchrun_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm must still use every core it may run on
    (round-1 SCALE ratios at N >= 2 were void because it ran on one)."""
    out = _run_bench(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "2000", "--gpus", "2"],
                     {"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2"})
    d = json.loads(out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["n_gpus"] == 2 and d["unit"] == "frames/s" and d["higher_is_better"] is True


def test_reference_arm_other_ranks_print_nothing():
    out = _run_bench(["--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert out.strip() == ""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "LDPC_128")), reason="reference tree not present (GPU box)")
def test_reference_arm_takes_the_tf_path_when_tensorflow_imports():
    """With an importable `tensorflow` (here: the NumPy shim, explicitly allowed) and a reference tree, the arm runs the
    unmodified reference instead of the C port, and labels what it ran."""
    out = _run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--order", "1"],
                     {"LDPCB_ALLOW_TF_SHIM": "1", "LDPCB_TF_FRAMES": "60", "LDPCB_REFERENCE_ROOT": REF,
                      "PYTHONPATH": os.path.join(ROOT, "oracle", "tf_shim")})
    d = json.loads(out.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"].startswith("tf")
    assert "unmodified reference" in d["config"]["note"] and d["value"] > 0
    # without the explicit permission the shim is refused and the C port runs
    out = _run_bench(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "1000"],
                     {"LDPCB_REFERENCE_ROOT": REF, "PYTHONPATH": os.path.join(ROOT, "oracle", "tf_shim")})
    d = json.loads(out.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and "shim" in d["config"]["note"]


def test_traffic_file_of_another_build_is_refused(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    f = tmp_path / "t.json"
    f.write_text(json.dumps({"build_stamp": "not-this-build", "nms_kernel": {"warp_instr_per_frame": 1, "dram_bytes_per_frame": 1}}))
    monkeypatch.setattr(bench, "TRAFFIC_FILE", str(f))
    prof, why = bench.load_traffic()
    assert prof == {} and "refused" in why
    from short_ldpc_decoding_osd_b200 import build as B

    f.write_text(json.dumps({"build_stamp": B._stamp(), "nms_kernel": {"warp_instr_per_frame": 7, "dram_bytes_per_frame": 512}}))
    prof, why = bench.load_traffic()
    assert prof["nms_kernel"]["warp_instr_per_frame"] == 7


class _FakeBlock:
    """Stands in for _lib._PinnedBlock (cudaMallocHost needs a driver): same ownership protocol over a bytearray."""
    freed = 0

    def __init__(self, nbytes):
        self._mem = np.zeros(max(nbytes, 1), np.uint8)
        self.__array_interface__ = {"data": (self._mem.ctypes.data, False), "shape": (max(nbytes, 1),), "typestr": "|u1", "version": 3}

    def __del__(self):
        type(self).freed += 1


def test_pinned_memory_outlives_the_array_it_was_sliced_from():
    """a view of pinned_empty()'s result must keep the block alive after the original is collected."""
    _FakeBlock.freed = 0
    blocks = []

    def alloc(n):
        b = _FakeBlock(n)
        blocks.append(weakref.ref(b))
        return b

    x = _lib.pinned_empty((64, 128), np.float32, _alloc=alloc)
    assert x.shape == (64, 128) and x.dtype == np.float32 and x.flags["C_CONTIGUOUS"] and x.flags["WRITEABLE"]
    view = x[:7]
    flat = x.reshape(-1)[5:9]
    del x
    gc.collect()
    assert blocks[0]() is not None and _FakeBlock.freed == 0, "block freed while views are alive"
    view[:] = 3.0
    flat[:] = 4.0
    assert float(view[6, 127]) == 3.0 and float(flat[0]) == 4.0
    del view
    gc.collect()
    assert _FakeBlock.freed == 0
    del flat
    gc.collect()
    assert blocks[0]() is None and _FakeBlock.freed == 1, "block must be freed with its last view"
    z = _lib.pinned_empty((0, 4), np.uint32, _alloc=alloc)
    assert z.shape == (0, 4)


def test_numa_helpers_do_not_need_a_gpu():
    assert _lib.device_numa_node(0) == -1  # no device here: unknown, never an exception
    info = _lib.bind_host_to_device(0)
    assert info["bound"] is False
