import sys, os, torch
sys.path.insert(0, '/root/repo')
import hybrid_als_twotower_recommender_b200
from hybrid_als_twotower_recommender_b200 import scoring
for U in (16384, 65536):
    I=1250000; ka,kt=128,50
    g = torch.Generator(device="cuda").manual_seed(1)
    Ua = torch.randn(U, ka, device="cuda", generator=g) * ka ** -0.5
    Ia = torch.randn(I, ka, device="cuda", generator=g)
    Ut = torch.nn.functional.layer_norm(torch.randn(U, kt, device="cuda", generator=g), (kt,))
    It = torch.nn.functional.layer_norm(torch.randn(I, kt, device="cuda", generator=g), (kt,))
    sc = scoring.HybridScorer(Ua, Ia, Ut, It)
    for r in range(4):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); ex = sc.extrema(); e1.record(); torch.cuda.synchronize()
        print(U, "extrema ms", e0.elapsed_time(e1), "flagged pass1", sc.flagged_users(U, 0))
    del sc, Ua, Ia, Ut, It
