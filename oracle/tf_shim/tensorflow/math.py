import numpy as np

from . import _c, _t


def mod(x, y):
    return _t(np.mod(np.asarray(_c(x)), np.asarray(_c(y))))


def exp(x):
    return _t(np.exp(np.asarray(_c(x))))


def abs(x):  # noqa: A001
    return _t(np.abs(np.asarray(_c(x))))


def not_equal(x, y):
    return _t(np.not_equal(np.asarray(_c(x)), np.asarray(_c(y))))


def erf(x):
    from scipy.special import erf as _erf

    a = np.asarray(_c(x))
    return _t(_erf(a).astype(a.dtype))
