"""CPU ORACLE (test infrastructure, NOT product code) -- ALS half-step / fit / predict.

Restates, in numpy fp64, the arithmetic that the reference reaches through
``pyspark.ml.recommendation.ALS`` at ``src/als_model.py:52-62`` (fit) and
``src/als_model.py:75`` (transform).  The algorithm itself lives in a
third-party dependency that is NOT under /root/reference:

    pyspark==3.5.1 (requirements.txt:1)
      -> org.apache.spark:spark-mllib_2.12:3.5.1
         mllib/src/main/scala/org/apache/spark/ml/recommendation/ALS.scala
         (NormalEquation.add / CholeskySolver.solve / computeYtY / computeFactors)

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path and neither pyspark nor a JVM exists in
this image, so this restatement cannot be checked against Spark output.  It is
anchored on the reference's call-site parameters (als_model.py:52-60: rank,
maxIter, regParam, coldStartStrategy; every other Spark parameter at its
default) and on the published Spark algorithm:

  * ids Int, ratings Float, factors Float; normal equations and solve in Double
  * sweep order: item half-step first, then user half-step
  * explicit:  A += y y^T (dspr, packed upper), b += r*y, n += 1
  * implicit:  A starts at Y^T Y; c1 = alpha*|r|; A += c1 * y y^T;
               if r > 0: b += (1 + c1)*y, n += 1
  * A[d,d] += regParam * n   (ALS-WR weighting, lambda * n_row, not lambda*I)
  * solve by packed Cholesky (LAPACK dppsv, uplo='U'); x cast to Float
  * duplicates (u,i) are not merged
  * predict: sequential Float dot; a side with no factor row -> NaN

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import numpy as np

try:  # LAPACK packed Cholesky, the routine Spark's CholeskySolver calls
    from scipy.linalg.lapack import dppsv as _dppsv
except Exception:  # pragma: no cover
    _dppsv = None


# --------------------------------------------------------------------------
# CSR helpers (host logic only; the product has its own GPU CSR builder)
# --------------------------------------------------------------------------
def coo_to_csr(rows, cols, vals, n_rows):
    """Stable COO->CSR: ratings of a row keep their input order (Spark does not
    merge duplicates; als_model.py:51 hands the raw frame to Spark)."""
    rows = np.asarray(rows, dtype=np.int64)
    order = np.argsort(rows, kind="stable")
    counts = np.bincount(rows, minlength=n_rows)
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, np.asarray(cols, dtype=np.int32)[order], np.asarray(vals, dtype=np.float32)[order]


def gram_f64(Y):
    """computeYtY: sum of y y^T over source rows, fp64."""
    Yd = np.asarray(Y, dtype=np.float64)
    return Yd.T @ Yd


def _packed_upper(A):
    k = A.shape[0]
    ap = np.empty(k * (k + 1) // 2, dtype=np.float64)
    p = 0
    for j in range(k):
        ap[p:p + j + 1] = A[: j + 1, j]
        p += j + 1
    return ap


def solve_packed_cholesky(A, b):
    """CholeskySolver.solve: dppsv('U') on the packed upper triangle."""
    if _dppsv is not None:
        x, info = _dppsv(A.shape[0], _packed_upper(A), b.reshape(-1, 1).copy(), lower=0)
        if info != 0:
            raise np.linalg.LinAlgError(f"dppsv info={info}")
        return np.asarray(x).reshape(-1)
    L = np.linalg.cholesky(A)
    return np.linalg.solve(L.T, np.linalg.solve(L, b))


# --------------------------------------------------------------------------
# Half-step
# --------------------------------------------------------------------------
def als_half_step_loops(rowptr, colidx, vals, src, reg, implicit=False, alpha=1.0):
    """Literal per-rating restatement (pure loops; small cases only).

    Follows NormalEquation.add rating by rating: fp64 rank-1 updates in CSR
    order, b update, n count, lambda*n on the diagonal, dppsv, cast to fp32.
    Rows with no stored rating keep a zero factor (they are absent from the
    Spark model; `present` mask is returned separately).
    """
    m = len(rowptr) - 1
    k = src.shape[1]
    out = np.zeros((m, k), dtype=np.float32)
    yty = gram_f64(src) if implicit else None
    for j in range(m):
        lo, hi = int(rowptr[j]), int(rowptr[j + 1])
        if hi == lo:
            continue
        A = yty.copy() if implicit else np.zeros((k, k), dtype=np.float64)
        b = np.zeros(k, dtype=np.float64)
        n = 0
        for p in range(lo, hi):
            y = src[colidx[p]].astype(np.float64)
            r = float(vals[p])
            if implicit:
                c1 = alpha * abs(r)
                A += c1 * np.outer(y, y)
                if r > 0.0:
                    b += (1.0 + c1) * y
                    n += 1
            else:
                A += np.outer(y, y)
                b += r * y
                n += 1
        A[np.diag_indices(k)] += reg * n
        out[j] = solve_packed_cholesky(A, b).astype(np.float32)
    return out


def als_half_step(rowptr, colidx, vals, src, reg, implicit=False, alpha=1.0):
    """Vectorised fp64 restatement of the same half-step (same maths as
    als_half_step_loops; summation order inside a row differs only at the
    1e-16 level).  Used for the mid-size parity cases."""
    rowptr = np.asarray(rowptr)
    m = len(rowptr) - 1
    k = src.shape[1]
    out = np.zeros((m, k), dtype=np.float32)
    srcd = np.asarray(src, dtype=np.float64)
    yty = srcd.T @ srcd if implicit else None
    eye = np.eye(k)
    for j in range(m):
        lo, hi = int(rowptr[j]), int(rowptr[j + 1])
        if hi == lo:
            continue
        G = srcd[colidx[lo:hi]]
        r = vals[lo:hi].astype(np.float64)
        if implicit:
            c1 = alpha * np.abs(r)
            A = yty + (G * c1[:, None]).T @ G
            pos = r > 0.0
            b = ((1.0 + c1) * pos) @ G
            n = int(pos.sum())
        else:
            A = G.T @ G
            b = r @ G
            n = hi - lo
        A = A + (reg * n) * eye
        out[j] = solve_packed_cholesky(A, b).astype(np.float32)
    return out


def row_present(rowptr):
    rowptr = np.asarray(rowptr)
    return (rowptr[1:] - rowptr[:-1]) > 0


# --------------------------------------------------------------------------
# Fit / predict / RMSE
# --------------------------------------------------------------------------
def als_fit(users, items, ratings, n_users, n_items, rank, max_iter, reg,
            init_user_factors, implicit=False, alpha=1.0, half_step=None):
    """ALS.train loop: item half-step FIRST, then user half-step, max_iter times.

    Spark's own initialisation (XORShiftRandom; pyspark's default seed depends on
    Python hash randomisation) is not reproducible, so initial user factors are an
    explicit argument; initial item factors are never read (item step runs first).
    Returns (user_factors fp32 [U,k], item_factors fp32 [I,k]).
    """
    hs = half_step or als_half_step
    ur, uc, uv = coo_to_csr(users, items, ratings, n_users)    # R   : user rows
    ir, ic, iv = coo_to_csr(items, users, ratings, n_items)    # R^T : item rows
    X = np.ascontiguousarray(init_user_factors, dtype=np.float32)
    assert X.shape == (n_users, rank)
    Y = np.zeros((n_items, rank), dtype=np.float32)
    for _ in range(max_iter):
        Y = hs(ir, ic, iv, X, reg, implicit, alpha)
        X = hs(ur, uc, uv, Y, reg, implicit, alpha)
    return X, Y


def als_predict(X, Y, users, items, user_present=None, item_present=None):
    """ALSModel.transform: sequential fp32 dot; missing side -> NaN
    (coldStartStrategy='drop' removes those rows afterwards, als_model.py:22,59)."""
    users = np.asarray(users)
    items = np.asarray(items)
    k = X.shape[1]
    acc = np.zeros(len(users), dtype=np.float32)
    xu = X[users]
    yi = Y[items]
    for f in range(k):
        acc = (acc + xu[:, f] * yi[:, f]).astype(np.float32)
    if user_present is not None:
        acc = np.where(user_present[users], acc, np.float32(np.nan))
    if item_present is not None:
        acc = np.where(item_present[items], acc, np.float32(np.nan))
    return acc


def rmse(X, Y, users, items, ratings):
    pred = als_predict(X, Y, users, items).astype(np.float64)
    return float(np.sqrt(np.mean((pred - np.asarray(ratings, dtype=np.float64)) ** 2)))


def init_factors(n, rank, seed):
    """The distribution Spark's `initialize` uses: i.i.d. N(0,1) rows scaled to
    unit L2 norm, fp32.  (Spark's generator is XORShiftRandom; only the
    distribution is reproduced, the harness feeds identical factors to oracle and
    kernels.)"""
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((n, rank)).astype(np.float32)
    nrm = np.sqrt((f.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    nrm[nrm == 0] = 1.0
    return (f / nrm[:, None]).astype(np.float32)
