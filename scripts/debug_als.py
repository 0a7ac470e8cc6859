import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_als import gpu_half_step, synth
from oracle import als_oracle, c_oracle
from tests.util import rel_l2
for k in [2, 4, 10, 16, 20, 32, 50, 64, 100, 128]:
    U, I, nnz = 700, 300, 9000
    u, i, r = synth(U, I, nnz, k)
    X = als_oracle.init_factors(U, k, 1)
    got, _ = gpu_half_step(i, u, r, I, X, 0.1)
    rp, ci, v = als_oracle.coo_to_csr(i, u, r, I)
    want = c_oracle.als_half_step(rp, ci, v, X, 0.1)
    cnt = np.diff(rp)
    err = np.abs(got - want).max(1)
    print(k, "rel", rel_l2(got, want), "maxabs", err.max(), "bad rows", int((err > 1e-3).sum()), "of", I,
          "len of bad", cnt[err > 1e-3][:8], "len of good", cnt[err <= 1e-3][:8])
