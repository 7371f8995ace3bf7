"""Drop-in for the reference's NMS test model (``LDPC_128/Ldpc_128_testing/ms_test.py``).

Same classes, methods, argument meaning and return shapes:

* ``Decoder_Layer(initial_value=-0.048)`` with the raw (pre-softplus) weights ``shared_check_weight``
  (and ``shared_bit_weight`` / ``shared_bit_weight1`` / ``shared_bit_weight2`` for NMS-2 / NMS-3)
  -- ms_test.py:71-97; ``layer(soft_input, labels)`` -> list of ``num_iterations + 1`` arrays
  ``float32[B,128]`` (index 0 = input) -- ms_test.py:99-121.
* ``Decoding_model()(inputs, labels)`` -> ``(fer, ber, undetected_count, (buffer_inputs, buffer_labels))``
  -- ms_test.py:30-34; ``get_eval`` -- :36-54; ``collect_failed_output_selective`` -- :55-64;
  ``postprocess_failure_cases`` -- :66-70.

All arithmetic runs in libldpc_b200.so (CUDA, sm_100a) through the host-buffer C-ABI calls; NumPy only
reshapes the results into the reference's Python structures.  Weights exported from a TF checkpoint
are assigned to the ``shared_*`` attributes as plain floats/arrays.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import _lib
from . import globalmap as GL
from .runtime import get_handle, softplus


def _as_llr(x) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x), dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 128:
        raise ValueError(f"expected float[B,128] soft input, got shape {a.shape}")
    return a


class Decoder_Layer:
    def __init__(self, initial_value: float = -0.048):
        self.decoder_type = GL.get_map("selected_decoder_type") or "NMS-1"
        self.num_iterations = int(GL.get_map("num_iterations") or 12)
        self.code = GL.get_map("code_parameters")
        self.initials = initial_value
        if self.decoder_type not in ("NMS-1", "NMS-2", "NMS-3"):
            raise NotImplementedError(f"decoder type {self.decoder_type!r}: only NMS-1/2/3 are on the hot path (ms_test.py:84-92)")
        w = lambda: np.full([1], initial_value, dtype=np.float32)  # noqa: E731
        self.shared_check_weight = w()
        if self.decoder_type == "NMS-2":
            self.shared_bit_weight = w()
        if self.decoder_type == "NMS-3":
            self.shared_bit_weight1 = w()
            self.shared_bit_weight2 = w()

    # (alpha, w_vc, w_marg) as the kernel takes them: softplus applied here (ms_test.py:127-131,207-208,222-226)
    def kernel_weights(self) -> Tuple[float, float, float]:
        alpha = softplus(np.asarray(self.shared_check_weight).reshape(-1)[0])
        w_vc = w_marg = 1.0
        if self.decoder_type == "NMS-2":
            w_vc = w_marg = softplus(np.asarray(self.shared_bit_weight).reshape(-1)[0])
        if self.decoder_type == "NMS-3":
            w_vc = softplus(np.asarray(self.shared_bit_weight1).reshape(-1)[0])
            w_marg = softplus(np.asarray(self.shared_bit_weight2).reshape(-1)[0])
        return alpha, w_vc, w_marg

    def __call__(self, soft_input, labels=None) -> List[np.ndarray]:
        return self.call(soft_input, labels)

    def call(self, soft_input, labels=None) -> List[np.ndarray]:
        y = _as_llr(soft_input)
        B = y.shape[0]
        alpha, w_vc, w_marg = self.kernel_weights()
        h = get_handle(self.code)
        bits = np.empty((B, 4), np.uint32)
        traj = np.empty((B, self.num_iterations + 1, 128), np.float32)
        h.call("ldpcb_nms_decode_host", y, B, self.num_iterations, alpha, w_vc, w_marg, 0, bits, None, None, traj)
        return [traj[:, i, :] for i in range(self.num_iterations + 1)]


class _Rows:
    """Read-only sequence of the rows of a 2-D array (optionally row i -> base[index[i]]): what the reference builds as
    a Python list of 13*F tensors (ms_test.py:55-64), without materialising 13*F objects.  len(), iteration, indexing
    and np.asarray() behave like the list of rows; `.array()` returns the stacked rows (a view when possible)."""

    def __init__(self, base: np.ndarray, index: np.ndarray = None):
        self._base, self._index = base, index

    def __len__(self):
        return len(self._base) if self._index is None else len(self._index)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self._base[i] if self._index is None else self._base[self._index[i]]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def array(self) -> np.ndarray:
        return self._base if self._index is None else self._base[self._index]

    def __array__(self, dtype=None, copy=None):
        a = self.array()
        return a if dtype is None else a.astype(dtype)

    def __add__(self, other):  # list concatenation, as the reference's drivers do with the per-batch buffers
        return list(self) + list(other)

    def __radd__(self, other):
        return list(other) + list(self)


def _pack_labels(lab: np.ndarray) -> np.ndarray:
    """int[B,128] 0/1 labels -> uint32[B,4] (one cast, one packbits)."""
    b = lab if lab.dtype == np.uint8 else lab.astype(np.uint8)
    return np.packbits(np.ascontiguousarray(b), axis=1, bitorder="little").view("<u4")


class Decoding_model:
    def __init__(self):
        self.layer = Decoder_Layer()

    def __call__(self, inputs, labels):
        return self.call(inputs, labels)

    def call(self, inputs, labels):
        """One C-ABI call (ldpcb_nms_retest_host): NMS with the get_eval tallies on the GPU, the detected failures
        compacted and re-decoded with trajectory output on the device, their 13 rows copied back -- the retest
        records of ms_test.py:55-64.  The buffer is returned as two row sequences backed by one array each."""
        y = _as_llr(inputs)
        lab = np.asarray(labels)
        B = y.shape[0]
        L = self.layer
        alpha, w_vc, w_marg = L.kernel_weights()
        h = get_handle(L.code)
        rows = L.num_iterations + 1
        truth = _pack_labels(lab)
        bits = np.empty((B, 4), np.uint32)
        syn = np.empty(B, np.uint8)
        cnt = np.zeros(_lib.NUM_COUNTERS, np.uint64)
        # room for the records: half the batch, or what the last call's failure rate suggests (low SNR)
        cap = max(64, int(B * max(0.5, min(1.0, 1.2 * getattr(self, "_fail_rate", 0.0)))))
        while True:
            fidx = np.empty(cap, np.int32)
            # pinned (recycled) memory for the 13-row records, most of the traffic of a call; very large batches take
            # pageable memory: pinning hundreds of MB per call costs more than it saves
            big = cap * rows * 512 > (64 << 20)
            traj = np.empty((cap, rows, 128), np.float32) if big else _lib.pinned_pool_empty((cap, rows, 128), np.float32)
            nf = np.zeros(1, np.int64)
            cnt[:] = 0
            h.call("ldpcb_nms_retest_host", y, B, L.num_iterations, alpha, w_vc, w_marg, truth, bits, syn, cnt, cap, fidx, traj, nf)
            n = int(nf[0])
            if n <= cap:
                break
            cap = n  # more failures than the buffer held (low SNR): once more with room for all of them
        self._fail_rate = n / max(B, 1)
        fer = float(cnt[1]) / B
        ber = float(cnt[2]) / (B * lab.shape[1])
        undetected = int(cnt[4])
        if undetected:
            und = np.flatnonzero((syn == 0) & (bits != truth).any(axis=1))
            print("Undetected Elements:", und[:, None])
        self.last_hard_bits = bits
        self.last_counters = cnt
        self.last_index = fidx[:n].astype(np.int64)[:, None]
        # the reference's layout (ms_test.py:60-63): for each failure its 13 rows, the label repeated 13 times
        buffer_inputs = _Rows(traj[:n].reshape(n * rows, 128))
        buffer_labels = _Rows(lab, np.repeat(fidx[:n].astype(np.int64), rows))
        return fer, ber, undetected, (buffer_inputs, buffer_labels)

    def get_eval(self, soft_output_list, labels):
        """(FER, BER, n_undetected, index int64[F,1]) from the last soft output (ms_test.py:36-54).
        The hard decision, syndrome and tallies are the kernel's (a zero-iteration decode of the given
        posterior is exactly tf.where(x>0,0,1) + syndrome)."""
        soft = _as_llr(soft_output_list[-1])
        lab = np.asarray(labels)
        B = soft.shape[0]
        h = get_handle(self.layer.code)
        truth = _lib.pack_bits(lab)
        bits = np.empty((B, 4), np.uint32)
        syn = np.empty(B, np.uint8)
        cnt = np.zeros(_lib.NUM_COUNTERS, np.uint64)
        h.call("ldpcb_decode_host", soft, B, 0, 1.0, 1.0, 1.0, 0, -1, 0, bits, syn, None, truth, cnt)
        if int(cnt[4]):
            und = np.flatnonzero((syn == 0) & (bits != truth).any(axis=1))
            print("Undetected Elements:", und[:, None])
        index = np.flatnonzero(syn)[:, None].astype(np.int64)
        return float(cnt[1]) / B, float(cnt[2]) / (B * lab.shape[1]), int(cnt[4]), index

    def collect_failed_output_selective(self, soft_output_list, labels, index):
        list_length = self.layer.num_iterations + 1
        buffer_inputs, buffer_labels = [], []
        for i in np.asarray(index).reshape(-1):
            for j in range(list_length):
                buffer_inputs.append(soft_output_list[j][i])
                buffer_labels.append(labels[i])
        return buffer_inputs, buffer_labels

    def postprocess_failure_cases(self, buffer):
        if len(buffer[0]) and all(isinstance(b, _Rows) for b in buffer[0]) and all(isinstance(b, _Rows) for b in buffer[1]):
            return _Rows(np.concatenate([b.array() for b in buffer[0]])), _Rows(np.concatenate([b.array() for b in buffer[1]]))
        buffer_inputs = [j for i in buffer[0] for j in i]
        buffer_labels = [j for i in buffer[1] for j in i]
        return buffer_inputs, buffer_labels


def save_decoded_data(updated_buffer, file_dir, snr, log_filename, list_length):
    """ms_test.py:251-272: per-iteration mean cross-entropy of the failed frames appended to the log, and the
    13-rows-per-failure retest file written as TFRecords (the input of PB_OSD / FS_OSD / DL_OSD_Testing_serial)."""
    from . import read_TFdata

    if len(updated_buffer[0]) == 0:
        info = np.zeros((0, 128), np.float32)
        label = np.zeros((0, 128), np.int64)
    elif isinstance(updated_buffer[0], _Rows):
        info = np.asarray(updated_buffer[0].array(), dtype=np.float32)
        label = np.asarray(updated_buffer[1].array()).astype(np.int64)
    else:
        info = np.stack([np.asarray(b, dtype=np.float32) for b in updated_buffer[0]])
        label = np.stack([np.asarray(b) for b in updated_buffer[1]]).astype(np.int64)
    CE_loss_list = []
    n_cases = 0
    for i in range(list_length):
        bits, labs = info[i::list_length], label[i::list_length]
        n_cases = bits.shape[0]
        CE_loss_list.append(calculation_loss(bits, labs) / max(n_cases, 1))
    print(CE_loss_list)
    with open(log_filename, "a+") as f:
        f.write(str(n_cases) + "tested:\n")
        f.write("# CE list:\n")
        f.write(" ".join(map(str, CE_loss_list)) + "\n")
    print("Data for retraining  with %d cases to be stored " % info.shape[0])
    read_TFdata.make_tfrecord((info, label), out_filename=file_dir)
    print("For " + str(round(snr, 2)) + "dB:Data storing finished!")
    return CE_loss_list


def calculation_loss(soft_output, labels) -> float:
    """Sum of sigmoid cross-entropies with logits = -soft_output (ms_test.py:244-249); log-only statistic."""
    x = -np.asarray(soft_output, dtype=np.float64)
    z = np.asarray(labels, dtype=np.float64)
    return float(np.sum(np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))))
