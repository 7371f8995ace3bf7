"""Drop-in for the two inference-time networks of ``LDPC_128/DL_OSD_Testing_serial/nn_net.py``.

* ``conv_bitwise`` (``:174-208``): per bit, the 13-sample LLR trajectory goes through Conv1D(8,k3) ->
  Conv1D(4,k3) -> Conv1D(2,k3) (all linear, no bias, 'valid') -> Flatten (7x2) -> Dense(1, bias).  Being fully
  linear it is exactly a 13-tap FIR plus bias per bit; ``set_weights`` folds exported Keras kernels into the taps
  once on the host and ``__call__`` runs ldpcb_dia_fir on the GPU.
* ``Predict_outlier_light`` (``:136-149``): Dense(6, no bias, linear) -> Dense(2, no bias, softmax) on the sorted
  window + position; six multiply-adds per decision, evaluated on the host inside the window policy.

Weights come from the reference's TF checkpoints as plain arrays (exporter: SURVEY.md 8f row f4).
"""
from __future__ import annotations

import numpy as np

from . import globalmap as GL
from .runtime import get_handle


def fold_conv_bitwise(k1, k2, k3, dense_w, dense_b, length=13):
    """Keras kernels k1[3,1,8], k2[3,8,4], k3[3,4,2], dense_w[14,1], dense_b[1] -> (taps[length], bias)."""
    k1, k2, k3 = (np.asarray(k, dtype=np.float64) for k in (k1, k2, k3))
    dw = np.asarray(dense_w, dtype=np.float64).reshape(-1)

    def conv(x, k):  # x [T, Cin], k [3, Cin, Cout], 'valid' cross-correlation as Keras Conv1D
        T = x.shape[0] - k.shape[0] + 1
        return np.stack([np.einsum("dc,dco->o", x[t:t + k.shape[0]], k) for t in range(T)], axis=0)

    def forward(x):
        return float(conv(conv(conv(x.reshape(-1, 1), k1), k2), k3).reshape(-1) @ dw)

    taps = np.array([forward(np.eye(length)[i]) for i in range(length)])
    return taps.astype(np.float32), float(np.asarray(dense_b).reshape(-1)[0])


class conv_bitwise:
    def __init__(self):
        code = GL.get_map("code_parameters")
        self.list_length = (GL.get_map("num_iterations") or 12) + 1
        self.n_dims = code.check_matrix_column if code is not None else 128
        self.taps = np.zeros(self.list_length, dtype=np.float32)
        self.taps[0] = 1.0  # identity on the channel LLR until weights are set
        self.bias = 0.0

    def set_weights(self, k1, k2, k3, dense_w, dense_b):
        self.taps, self.bias = fold_conv_bitwise(k1, k2, k3, dense_w, dense_b, self.list_length)

    def set_taps(self, taps, bias=0.0):
        self.taps = np.asarray(taps, dtype=np.float32).reshape(self.list_length)
        self.bias = float(bias)

    # nn_net.py:198-208
    def preprocessing_inputs(self, input_slice):
        original_input = np.asarray(input_slice[0], dtype=np.float32)
        original_label = np.asarray(input_slice[1])
        file_input_data = original_input.reshape(-1, self.list_length, self.n_dims)
        squashed_inputs = np.transpose(file_input_data, (0, 2, 1)).reshape(-1, self.list_length, 1)
        return squashed_inputs, original_input[0::self.list_length], original_label[0::self.list_length]

    def __call__(self, inputs):
        """squashed_inputs [B*128,13,1] -> float32[B,128] (nn_net.py:190-197)."""
        import torch  # device memory carrier only

        x = np.asarray(inputs, dtype=np.float32).reshape(-1, self.n_dims, self.list_length)
        traj = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))  # [B,13,128]
        B = traj.shape[0]
        h = get_handle()
        dev = f"cuda:{h.device}"
        td = torch.from_numpy(traj).to(dev)
        out = torch.empty((B, self.n_dims), dtype=torch.float32, device=dev)
        h.call("ldpcb_dia_fir", td, B, self.list_length, np.ascontiguousarray(self.taps), float(self.bias), out, None)
        torch.cuda.synchronize()
        return out.cpu().numpy()


class Predict_outlier_light:
    def __init__(self, sliding_win_width, W1=None, W2=None):
        self.input_width = sliding_win_width + 1
        self.W1 = np.eye(self.input_width, dtype=np.float32) if W1 is None else np.asarray(W1, dtype=np.float32)
        self.W2 = np.zeros((self.input_width, 2), dtype=np.float32) if W2 is None else np.asarray(W2, dtype=np.float32)

    def __call__(self, inputs):
        o = (np.asarray(inputs, dtype=np.float32) @ self.W1) @ self.W2
        e = np.exp(o - o.max(axis=-1, keepdims=True))
        return e / e.sum(axis=-1, keepdims=True)
