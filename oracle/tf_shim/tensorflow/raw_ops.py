import numpy as np

from . import _c, _t


def UniqueV2(x, axis):
    a = np.asarray(_c(x))
    seen, rows, idx = {}, [], []
    for r in a:  # first-occurrence order, like tf.unique
        key = tuple(np.atleast_1d(r).tolist())
        if key not in seen:
            seen[key] = len(rows)
            rows.append(r)
        idx.append(seen[key])
    return _t(np.array(rows)), _t(np.array(idx, dtype=np.int32))
