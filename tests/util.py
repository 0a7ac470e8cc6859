import numpy as np


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def check_topk_against_dense(B, idx, score, k, tol, item_offset=0):
    """B: oracle blend [U,I] fp64.  idx/score: [U,k] from the kernels.  Top-k index sets must be
    identical except for items whose oracle score lies within `tol` of the k-th score (ties)."""
    U, I = B.shape
    kk = min(k, I)
    for u in range(U):
        got = idx[u]
        valid = got[got >= 0] - item_offset
        assert len(valid) == kk, (u, len(valid), kk)
        assert len(set(valid.tolist())) == kk
        assert (got[kk:] == -1).all()
        s = score[u, :kk].astype(np.float64)
        assert np.all(np.diff(s) <= 0), "scores must be sorted descending"
        assert np.allclose(s, B[u, valid], rtol=0, atol=tol), np.abs(s - B[u, valid]).max()
        order = np.sort(B[u])[::-1]
        kth = order[kk - 1]
        must = set(np.nonzero(B[u] > kth + tol)[0].tolist())
        assert must <= set(valid.tolist()), (u, must - set(valid.tolist()))
        assert np.all(B[u, valid] >= kth - tol)
        # equal scores: lower item index first
        for a in range(kk - 1):
            if score[u, a] == score[u, a + 1]:
                assert got[a] < got[a + 1]
