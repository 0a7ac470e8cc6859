import sys, time, cProfile, pstats, numpy as np, torch, pandas as pd
sys.path.insert(0, '.')
import bench
from hybrid_als_twotower_recommender_b200 import ALSModel
w = bench.WORKLOADS["c2"]
u, i, r = bench.synth_coo(w, "cuda")
df = pd.DataFrame({"userId": u.cpu().numpy().astype(np.int64) * 7 + 3, "itemId": i.cpu().numpy().astype(np.int64) * 5 + 1,
                   "average_review_rating": r.cpu().numpy().astype(np.float64)})
m = ALSModel(rank=64, max_iter=10, reg_param=0.1)
m.train(df.iloc[:200000])
m.train(df)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); m.train(df); X = m.model.user_factors.cpu(); Y = m.model.item_factors.cpu(); t1 = time.perf_counter()
pr.disable()
print("ms", (t1 - t0) * 1e3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
