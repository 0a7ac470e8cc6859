import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_als_twotower_recommender_b200
from hybrid_als_twotower_recommender_b200 import scoring, _native as nat
U, I, ka, kt = 4096, 400000, 128, 50
g = torch.Generator(device="cuda").manual_seed(1)
Ua = torch.randn(U, ka, device="cuda", generator=g) * ka ** -0.5
Ia = torch.randn(I, ka, device="cuda", generator=g)
Ut = torch.nn.functional.layer_norm(torch.randn(U, kt, device="cuda", generator=g), (kt,))
It = torch.nn.functional.layer_norm(torch.randn(I, kt, device="cuda", generator=g), (kt,))
sc = scoring.HybridScorer(Ua, Ia, Ut, It)
ex = sc.extrema(); torch.cuda.synchronize()
print("flagged pass1:", sc.flagged_users(U, 0), "of", U)
Sa = Ua[:8].double() @ Ia.double().T; St = Ut[:8].double() @ It.double().T
print("exact  ", Sa.min(1).values[:3].tolist(), Sa.max(1).values[:3].tolist(), St.max(1).values[:3].tolist())
print("kernel ", ex[:3].tolist())
srt = torch.sort(Sa, dim=1, descending=True).values
print("top1-top4 gap (a):", (srt[:, 0] - srt[:, 3]).tolist()[:4], "|u_a|", Ua[:4].norm(dim=1).tolist(), "max|i_a|", float(Ia.norm(dim=1).max()))
print("eps_a:", (1.05 * 2 ** -8 * Ua[:4].norm(dim=1) * Ia.norm(dim=1).max()).tolist())
