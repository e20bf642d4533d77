"""ctypes loader for oracle/libgn_oracle.so (topk_oracle.c).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgn_oracle.so")


def available() -> bool:
    return os.path.exists(_PATH)


def _lib():
    lib = C.CDLL(_PATH)
    fp = C.POINTER(C.c_float)
    lib.gn_oracle_corr.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
    lib.gn_oracle_topk_h.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
    return lib


def corr(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    b, n, d = x.shape
    out = np.empty((b, n, n), dtype=np.float32)
    fp = C.POINTER(C.c_float)
    _lib().gn_oracle_corr(x.ctypes.data_as(fp), b, n, d, out.ctypes.data_as(fp))
    return out


def topk_h(corr_: np.ndarray, scale: int) -> np.ndarray:
    corr_ = np.ascontiguousarray(corr_, dtype=np.float32)
    b, n, _ = corr_.shape
    e = 1 if scale == n else n
    out = np.empty((b, e, n), dtype=np.float32)
    fp = C.POINTER(C.c_float)
    rc = _lib().gn_oracle_topk_h(corr_.ctypes.data_as(fp), b, n, int(scale), out.ctypes.data_as(fp))
    if rc == -3:
        raise RuntimeError("selected index k out of range")
    return out
