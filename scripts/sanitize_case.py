"""Small cases for compute-sanitizer (racecheck / memcheck): both tensor-core ALS ranks, long-row slices,
and the tensor-core scoring path with its exact re-run."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_als import gpu_half_step, synth
from oracle import als_oracle, c_oracle
which = sys.argv[1] if len(sys.argv) > 1 else "als"
if which == "als":
    for k, seg in ((64, 64), (128, 96)):
        U, I, nnz = 400, 24, 3000
        u, i, r = synth(U, I, nnz, 5, skew=True)
        X = als_oracle.init_factors(U, k, 2)
        got, plan = gpu_half_step(i, u, r, I, X, 0.1, seg_len=seg)
        rp, ci, v = als_oracle.coo_to_csr(i, u, r, I)
        want = c_oracle.als_half_step(rp, ci, v, X, 0.1)
        print(k, "long rows", plan.n_long, "maxabs", float(np.abs(got - want).max()))
else:
    import hybrid_als_twotower_recommender_b200
    from hybrid_als_twotower_recommender_b200 import scoring
    rng = np.random.default_rng(0)
    U, I, ka, kt, k = 130, 33000, 20, 8, 10
    Ua, Ia = rng.normal(size=(U, ka)).astype(np.float32), rng.normal(size=(I, ka)).astype(np.float32)
    Ut, It = rng.normal(size=(U, kt)).astype(np.float32), rng.normal(size=(I, kt)).astype(np.float32)
    sc = scoring.HybridScorer(*(torch.from_numpy(a).cuda() for a in (Ua, Ia, Ut, It)))
    idx, s = sc.recommend(k, 0.8, 0.2)
    torch.cuda.synchronize()
    print("scoring ok", idx.shape, sc.flagged_users(U, k))
