"""Tests that need a REAL TensorFlow and a copy of the reference tree.

They are collected everywhere and skip cleanly where either is missing (this container and the GPU box have no
TensorFlow and no network); on an image with TensorFlow they check, with no further work:
  * the TF-free TFRecord reader against a file written by the reference's own writer
    (LDPC_128/Ldpc_128_testing/data_generating.py:8-26), and the reverse: tf.data parsing a file our writer wrote
    with the reference's reader (read_TFdata.py:10-29);
  * scripts/export_tf_weights.py on a checkpoint written the way the reference writes it
    (tf.train.Checkpoint(myAwesomeModel=model), Ldpc_128_training/training_stage.py:24, ldpc_128_testing.py:57-68);
  * the NumPy TensorFlow shim that pins the goldens (oracle/tf_shim) against TensorFlow itself: every
    tests/golden/*_ref_shim.npz fixture regenerated under real TF must equal the committed one.
"""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in (os.environ.get("LDPCB_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "LDPC_128", "Ldpc_128_testing")):
            return cand
    return None


def _real_tf():
    spec = importlib.util.find_spec("tensorflow")
    if spec is None or "tf_shim" in (spec.origin or ""):
        return None
    try:
        import tensorflow as tf
    except Exception:
        return None
    return tf


tf = _real_tf()
REF = _reference_root()
needs_tf = pytest.mark.skipif(tf is None, reason="TensorFlow is not installed (offline image)")
needs_ref = pytest.mark.skipif(REF is None, reason="no copy of the reference tree (baseline/_ref, $LDPCB_REFERENCE_ROOT)")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@needs_tf
@needs_ref
def test_reader_reads_a_record_file_written_by_the_reference(tmp_path):
    from short_ldpc_decoding_osd_b200 import read_TFdata as ours

    ref_writer = _load(os.path.join(REF, "LDPC_128", "Ldpc_128_testing", "data_generating.py"), "ref_data_generating")
    rng = np.random.default_rng(0)
    feats = rng.normal(size=(37, 128)).astype(np.float32)
    labels = rng.integers(0, 2, (37, 128)).astype(np.int64)
    f = str(tmp_path / "ref.tfrecord")
    ref_writer.make_tfrecord((feats, labels), f)
    got = list(ours.data_handler(128, f, batch_size=37).as_numpy_iterator())
    assert len(got) == 1
    assert np.array_equal(got[0][0], feats) and np.array_equal(got[0][1], labels)


@needs_tf
@needs_ref
def test_reference_reader_reads_a_record_file_we_wrote(tmp_path, monkeypatch):
    from short_ldpc_decoding_osd_b200 import read_TFdata as ours

    rng = np.random.default_rng(1)
    feats = rng.normal(size=(13 * 5, 128)).astype(np.float32)        # the retest layout: 13 rows per failure
    labels = np.repeat(rng.integers(0, 2, (5, 128)), 13, axis=0).astype(np.int64)
    f = str(tmp_path / "ours.tfrecord")
    ours.make_tfrecord((feats, labels), f)
    d = os.path.join(REF, "LDPC_128", "Ldpc_128_testing")
    monkeypatch.syspath_prepend(d)
    ref_reader = _load(os.path.join(d, "read_TFdata.py"), "ref_read_TFdata")
    ds = ref_reader.data_handler(128, f, 13)
    got = list(ds.as_numpy_iterator())
    assert len(got) == 5
    assert np.array_equal(np.concatenate([g[0] for g in got]), feats)
    assert np.array_equal(np.concatenate([g[1] for g in got]), labels)


@needs_tf
@needs_ref
def test_exporter_reads_a_checkpoint_written_like_the_reference(tmp_path):
    """Write NMS, CNN and fcn checkpoints through the reference's own model classes with
    tf.train.Checkpoint(myAwesomeModel=...) + CheckpointManager, run the exporter as a user would, load the result."""
    from short_ldpc_decoding_osd_b200 import weights

    work = os.path.join(REF, "LDPC_128", "DL_OSD_Testing_serial")
    script = f"""
import sys, os
sys.path.insert(0, {work!r}); os.chdir({work!r})
import numpy as np, tensorflow as tf
import globalmap as GL, fill_matrix_info as F
GL.global_setting("python 2.0 3.0 6 100  12 CCSDS_ldpc_n128_k64.alist NMS-1".split())
code = F.Code(GL.get_map("H_filename")); GL.set_map("code_parameters", code)
import nn_net as N
out = {str(tmp_path)!r}
nn = N.conv_bitwise(); nn(tf.zeros([128, 13, 1]))
for v in nn.trainable_variables: v.assign(tf.random.stateless_normal(v.shape, seed=[1, len(v.shape)]))
tf.train.CheckpointManager(tf.train.Checkpoint(myAwesomeModel=nn), os.path.join(out, "cnn"), max_to_keep=5).save()
w = GL.get_map("sliding_win_width")
fcn = N.Predict_outlier_light(w); fcn(tf.zeros([1, w + 1]))
for v in fcn.trainable_variables: v.assign(tf.random.stateless_normal(v.shape, seed=[2, len(v.shape)]))
tf.train.CheckpointManager(tf.train.Checkpoint(myAwesomeModel=fcn), os.path.join(out, "fcn"), max_to_keep=5).save()
np.savez(os.path.join(out, "truth.npz"), k1=nn.cnv_one.kernel.numpy(), dense_b=nn.dense.bias.numpy(), fcn2=fcn.dense2.kernel.numpy())
"""
    subprocess.run([sys.executable, "-c", script], check=True, timeout=600)
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "export_tf_weights.py"), "--out", str(tmp_path / "weights.npz"),
                    "--cnn-ckpt", str(tmp_path / "cnn"), "--fcn-ckpt", str(tmp_path / "fcn")], check=True, cwd=work, timeout=600)
    w = weights.load_npz(str(tmp_path / "weights.npz"))
    truth = np.load(tmp_path / "truth.npz")
    for k in ("k1", "dense_b", "fcn2"):
        assert np.array_equal(w[k], truth[k])
    cnn = weights.make_conv_bitwise(w)
    assert np.asarray(cnn.taps).shape == (13,)


@needs_tf
@needs_ref
@pytest.mark.parametrize("which", ["nms", "osd", "fs", "pb", "dl"])
def test_goldens_regenerated_under_real_tensorflow_equal_the_shim_goldens(which, tmp_path):
    """oracle/ref_runner.py --real-tf runs the same unmodified reference files with TensorFlow instead of the NumPy
    shim and writes the fixture to another directory; every array must equal the committed (shim) fixture, float
    trajectories within the 1e-5 the north star allows and everything integer exactly."""
    env = dict(os.environ, LDPCB_REFERENCE_ROOT=REF)
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), which, "--real-tf", "--out", str(tmp_path)], check=True, env=env, timeout=3600)
    name = {"nms": "nms_ref_shim.npz", "osd": "osd_ref_shim.npz", "fs": "fs_ref_shim.npz", "pb": "pb_ref_shim.npz", "dl": "dl_ref_shim.npz"}[which]
    new, old = np.load(tmp_path / name), np.load(os.path.join(ROOT, "tests", "golden", name))
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        a, b = new[k], old[k]
        if a.dtype.kind == "f":
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-30, err_msg=f"{name}:{k}")
        else:
            assert np.array_equal(a, b), f"{name}:{k} differs between TensorFlow and the shim"


def test_these_tests_are_collected_and_skip_cleanly_offline():
    """Documents the state of THIS image; fails loudly if TensorFlow ever appears without the tests above running."""
    if tf is None:
        assert importlib.util.find_spec("tensorflow") is None or "tf_shim" in (importlib.util.find_spec("tensorflow").origin or "")
    else:
        assert hasattr(tf, "train") and hasattr(tf.train, "Checkpoint")
