import numpy as np

from . import _c, _t


def softplus(x):
    a = np.asarray(_c(x))
    return _t(np.log1p(np.exp(a)).astype(a.dtype))


def top_k(x, k=1, sorted=True):  # noqa: A002
    a = np.asarray(_c(x))
    idx = np.argsort(-a, axis=-1, kind="stable")[..., :k]
    return (_t(np.take_along_axis(a, idx, axis=-1)), _t(idx.astype(np.int32)))


def sigmoid_cross_entropy_with_logits(labels=None, logits=None):
    z, x = np.asarray(_c(labels)), np.asarray(_c(logits))
    return _t(np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x))))
