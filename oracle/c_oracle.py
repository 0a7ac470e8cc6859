"""ctypes binding of oracle/als_oracle.c (CPU ORACLE -- test infrastructure only).

Multi-threaded (OpenMP) restatement used for the larger parity cases and as the
`cpu_baseline` / `--impl reference` leg of bench.py ("port": Spark local[N] cannot
run in this image, see oracle/als_oracle.py header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libals_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "als_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, i32, f64, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p
        L.oracle_als_half_step.restype = i64
        L.oracle_als_half_step.argtypes = [vp, vp, vp, i64, i64, vp, i64, i32, f64, i32, f64, vp]
        L.oracle_gram_packed.restype = None
        L.oracle_gram_packed.argtypes = [vp, i64, i32, vp]
        L.oracle_als_predict.restype = None
        L.oracle_als_predict.argtypes = [vp, vp, i32, vp, vp, i64, vp]
        L.oracle_hybrid_topk.restype = None
        L.oracle_hybrid_topk.argtypes = [vp, vp, i32, vp, vp, i32, i64, i64, f64, f64, i32, vp, vp]
        L.oracle_num_threads.restype = i32
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [i32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n: int):
    """OpenMP threads of the oracle port (torch.distributed.run exports OMP_NUM_THREADS=1 to its children)."""
    lib().oracle_set_num_threads(int(n))


def als_half_step(rowptr, colidx, vals, src, reg, implicit=False, alpha=1.0,
                  row_begin=0, row_end=None, out=None):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    src = np.ascontiguousarray(src, dtype=np.float32)
    m = len(rowptr) - 1
    k = src.shape[1]
    if row_end is None:
        row_end = m
    if out is None:
        out = np.zeros((m, k), dtype=np.float32)
    bad = lib().oracle_als_half_step(_p(rowptr), _p(colidx), _p(vals), row_begin, row_end, _p(src),
                                     src.shape[0], k, float(reg), int(bool(implicit)), float(alpha), _p(out))
    if bad:
        raise np.linalg.LinAlgError(f"row {bad - 1}: normal equations not positive definite")
    return out


def gram(src):
    src = np.ascontiguousarray(src, dtype=np.float32)
    k = src.shape[1]
    ap = np.zeros(k * (k + 1) // 2, dtype=np.float64)
    lib().oracle_gram_packed(_p(src), src.shape[0], k, _p(ap))
    G = np.zeros((k, k))
    iu = np.triu_indices(k)
    # packed upper is column-major over the upper triangle
    p = 0
    for j in range(k):
        G[: j + 1, j] = ap[p:p + j + 1]
        p += j + 1
    return np.triu(G) + np.triu(G, 1).T


def als_predict(X, Y, users, items):
    X = np.ascontiguousarray(X, dtype=np.float32)
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    users = np.ascontiguousarray(users, dtype=np.int32)
    items = np.ascontiguousarray(items, dtype=np.int32)
    out = np.empty(len(users), dtype=np.float32)
    lib().oracle_als_predict(_p(X), _p(Y), X.shape[1], _p(users), _p(items), len(users), _p(out))
    return out


def hybrid_topk(Ua, Ia, Ut, It, w_als, w_tt, topk):
    Ua, Ia, Ut, It = (np.ascontiguousarray(a, dtype=np.float32) for a in (Ua, Ia, Ut, It))
    nu, ni = Ua.shape[0], Ia.shape[0]
    idx = np.empty((nu, topk), dtype=np.int32)
    sc = np.empty((nu, topk), dtype=np.float64)
    lib().oracle_hybrid_topk(_p(Ua), _p(Ia), Ua.shape[1], _p(Ut), _p(It), Ut.shape[1], nu, ni,
                             float(w_als), float(w_tt), int(topk), _p(idx), _p(sc))
    return idx, sc
