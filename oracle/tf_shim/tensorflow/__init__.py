"""NumPy emulation of the TensorFlow ops the reference's hot-path modules call (TEST INFRASTRUCTURE).

TensorFlow is not installable in the build container (no network), so the reference's own source
files (LDPC_128/Ldpc_128_testing/ms_test.py, PB_OSD/pb_testing.py, FS_OSD/convention_osd.py,
FS_OSD/fs_testing.py, DL_OSD_Testing_serial/ordered_statistics_decoding.py) are imported UNMODIFIED
from /root/reference with this package first on sys.path, and their outputs on seeded inputs are
committed as golden vectors (oracle/ref_runner.py -> tests/golden/*.npz).

Only eager-mode semantics that those files rely on are modelled:
  * dtypes: Python float -> float32, Python int -> int32 (tf.convert_to_tensor defaults)
  * tf.argsort: stable, lower index first on ties in BOTH directions (TF implements argsort with
    top_k on the values or their negation)
  * tf.nn.top_k: values descending; tf.argmin/argmax: first occurrence; tf.sign(0) = 0
  * tf.where(c, x, y) elementwise / tf.where(c) -> int64 coordinates
  * fp32 reductions: summed by NumPy in fp32 (TF leaves the order unspecified)
"""
import numpy as np

float32, float64, int32, int64, bool = np.float32, np.float64, np.int32, np.int64, np.bool_
newaxis = None


class Tensor(np.ndarray):
    def numpy(self):
        v = np.asarray(self)
        return v[()] if v.ndim == 0 else v

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        return np.asarray(arr).view(Tensor)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        # TF binary ops convert the non-tensor operand with the tensor's dtype as hint
        # (x * H with H an int64 NumPy array stays float32), unlike NumPy's promotion.
        tdt = next(i.dtype for i in inputs if isinstance(i, Tensor))
        conv = []
        for i in inputs:
            if isinstance(i, Tensor):
                conv.append(np.asarray(i))
            else:
                a = np.asarray(i)
                if ufunc.nin == 2 and a.dtype != tdt and a.dtype.kind in "iufb" and tdt.kind in "iuf":
                    a = a.astype(tdt)
                conv.append(a)
        if "out" in kwargs:
            kwargs["out"] = tuple(np.asarray(o) for o in kwargs["out"])
        out = getattr(ufunc, method)(*conv, **kwargs)
        if isinstance(out, tuple):
            return tuple(np.asarray(o).view(Tensor) for o in out)
        return np.asarray(out).view(Tensor)

    def __bool__(self):
        a = np.asarray(self)
        if a.size == 0:
            return False
        return np.ndarray.__bool__(a)

    def __hash__(self):
        return id(self)

    def __len__(self):
        if self.ndim == 0:
            raise TypeError("Scalar tensor has no len()")
        return self.shape[0]

    def __getitem__(self, item):
        return np.asarray(np.ndarray.__getitem__(np.asarray(self), _plain(item))).view(Tensor)


def _plain(item):
    if isinstance(item, Tensor):
        return np.asarray(item)
    if isinstance(item, tuple):
        return tuple(_plain(i) for i in item)
    return item


def _t(x):
    return np.asarray(x).view(Tensor)


def convert_to_tensor(x, dtype=None):
    if isinstance(x, np.ndarray):
        a = np.asarray(x)
        return _t(a.astype(dtype) if dtype is not None and a.dtype != dtype else a)
    if isinstance(x, (list, tuple)):
        a = np.array([np.asarray(e) for e in x]) if len(x) and isinstance(x[0], np.ndarray) else np.array(x)
    else:
        a = np.array(x)
    if dtype is not None:
        return _t(a.astype(dtype))
    if a.dtype == np.float64 and not isinstance(x, np.generic):
        a = a.astype(np.float32)
    elif a.dtype == np.int64 and not isinstance(x, np.generic) and not _has_np(x):
        a = a.astype(np.int32)
    return _t(a)


def _has_np(x):
    if isinstance(x, (np.ndarray, np.generic)):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_np(e) for e in x)
    return False


_c = convert_to_tensor


def constant(value, dtype=None, shape=None):
    t = _c(value, dtype)
    return reshape(t, shape) if shape is not None else t


def cast(x, dtype):
    return _t(np.asarray(_c(x)).astype(dtype))


def zeros(shape, dtype=float32):
    return _t(np.zeros(tuple(shape) if not np.isscalar(shape) else (shape,), dtype=dtype))


def ones(shape, dtype=float32):
    return _t(np.ones(tuple(shape) if not np.isscalar(shape) else (shape,), dtype=dtype))


def ones_like(x, dtype=None):
    return _t(np.ones_like(np.asarray(_c(x)), dtype=dtype))


def eye(n, dtype=float32):
    return _t(np.eye(n, dtype=dtype))


def range(*args, dtype=None):  # noqa: A001
    return _t(np.arange(*args, dtype=dtype or np.int32))


def size(x):
    return _t(np.int32(np.asarray(_c(x)).size))


def reshape(x, shape):
    return _t(np.reshape(np.asarray(_c(x)), tuple(int(s) for s in np.atleast_1d(np.asarray(shape)))))


def expand_dims(x, axis):
    return _t(np.expand_dims(np.asarray(_c(x)), axis))


def squeeze(x, axis=None):
    return _t(np.squeeze(np.asarray(_c(x)), axis=axis))


def transpose(x, perm=None):
    return _t(np.transpose(np.asarray(_c(x)), perm))


def tile(x, multiples):
    return _t(np.tile(np.asarray(_c(x)), tuple(multiples)))


def concat(values, axis):
    arrs = [np.asarray(v) for v in values]
    first = next((v for v in values if isinstance(v, Tensor)), None)
    if first is not None:
        arrs = [a.astype(first.dtype) for a in arrs]
    return _t(np.concatenate(arrs, axis=axis))


def stack(values, axis=0):
    return _t(np.stack([np.asarray(_c(v)) for v in values], axis=axis))


def where(condition, x=None, y=None):
    c = np.asarray(_c(condition)).astype(np.bool_)
    if x is None:
        return _t(np.argwhere(c).astype(np.int64))
    xa, ya = _c(x), _c(y)
    if xa.dtype != ya.dtype:  # TF requires equal dtypes; Python scalars adopt the tensor's dtype
        if not isinstance(x, (np.ndarray, np.generic)):
            xa = _c(x, ya.dtype)
        else:
            ya = _c(y, xa.dtype)
    return _t(np.where(c, np.asarray(xa), np.asarray(ya)))


def gather(params, indices, axis=None, batch_dims=0):
    p = np.asarray(_c(params))
    i = np.asarray(_c(indices)).astype(np.int64)
    if batch_dims == 0:
        return _t(np.take(p, i, axis=0 if axis is None else axis))
    assert batch_dims == 1 and (axis is None or axis == 1)
    return _t(np.stack([np.take(p[b], i[b], axis=0) for b in np.arange(p.shape[0])], axis=0))


def boolean_mask(tensor, mask, axis=None):
    t = np.asarray(_c(tensor))
    m = np.asarray(_c(mask)).astype(np.bool_)
    ax = 0 if axis is None else axis
    if m.ndim == 1:
        return _t(np.compress(m, t, axis=ax))
    assert ax == 0
    return _t(t[m])


def tensor_scatter_nd_update(tensor, indices, updates):
    out = np.array(np.asarray(_c(tensor)))
    idx = np.asarray(_c(indices)).astype(np.int64)
    out[tuple(idx.T)] = np.asarray(_c(updates))
    return _t(out)


def matmul(a, b, transpose_a=False, transpose_b=False):
    A, Bm = np.asarray(_c(a)), np.asarray(_c(b))
    if transpose_a:
        A = np.swapaxes(A, -1, -2)
    if transpose_b:
        Bm = np.swapaxes(Bm, -1, -2)
    return _t(np.matmul(A, Bm))


def _red(fn):
    def f(x, axis=None, keepdims=False):
        a = np.asarray(_c(x))
        if isinstance(axis, list):
            axis = tuple(axis)
        return _t(fn(a, axis=axis, keepdims=keepdims))
    return f


reduce_sum = _red(lambda a, axis, keepdims: np.sum(a, axis=axis, keepdims=keepdims, dtype=a.dtype if a.dtype != np.bool_ else None))
reduce_prod = _red(lambda a, axis, keepdims: np.prod(a, axis=axis, keepdims=keepdims, dtype=a.dtype))
reduce_min = _red(np.min)
reduce_max = _red(np.max)
reduce_mean = _red(lambda a, axis, keepdims: np.mean(a, axis=axis, keepdims=keepdims, dtype=a.dtype))
reduce_all = _red(np.all)
reduce_any = _red(np.any)


def argmin(x, axis=0, output_type=int64):
    return _t(np.argmin(np.asarray(_c(x)), axis=axis).astype(output_type))


def argmax(x, axis=0, output_type=int64):
    return _t(np.argmax(np.asarray(_c(x)), axis=axis).astype(output_type))


def argsort(values, axis=-1, direction="ASCENDING", stable=False, name=None):
    v = np.asarray(_c(values))
    key = v if direction == "ASCENDING" else -v.astype(np.float64 if v.dtype.kind == "f" else np.int64)
    return _t(np.argsort(key, axis=axis, kind="stable").astype(np.int32))


def sort(values, axis=-1, direction="ASCENDING", name=None):
    v = np.sort(np.asarray(_c(values)), axis=axis, kind="stable")
    return _t(v if direction == "ASCENDING" else np.flip(v, axis=axis))


def equal(x, y):
    return _t(np.equal(np.asarray(_c(x)), np.asarray(_c(y))))


def sign(x):
    return _t(np.sign(np.asarray(_c(x))))


def abs(x):  # noqa: A001
    return _t(np.abs(np.asarray(_c(x))))


def sqrt(x):
    return _t(np.sqrt(np.asarray(_c(x))))


def floor(x):
    return _t(np.floor(np.asarray(_c(x))))


def sigmoid(x):
    a = np.asarray(_c(x))
    return _t((1 / (1 + np.exp(-a))).astype(a.dtype))


def clip_by_value(t, lo, hi):
    a = np.asarray(_c(t))
    return _t(np.clip(a, a.dtype.type(lo), a.dtype.type(hi)))


def stop_gradient(x):
    return x


def while_loop(cond, body, loop_vars):
    v = list(loop_vars)
    while cond(*v):
        v = list(body(*v))
    return v


def print(*args, **kwargs):  # noqa: A001
    import builtins

    builtins.print(*args)


from . import keras, math, nn, random, raw_ops, train  # noqa: E402,F401
