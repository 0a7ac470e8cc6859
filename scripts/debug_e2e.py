"""Where the end-to-end call spends its time (c2 shape): H2D, CSR builds, plans, sweeps, D2H."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_als_twotower_recommender_b200
from hybrid_als_twotower_recommender_b200 import csr, als_engine
import bench
w = bench.WORKLOADS["c2"]
u, i, r = bench.synth_coo(w, torch.device("cuda"))
hu, hi, hr = (t.cpu().pin_memory() for t in (u, i, r))
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T()
    du, di, dr = hu.cuda(non_blocking=True), hi.cuda(non_blocking=True), hr.cuda(non_blocking=True)
    t1 = T()
    ucnt = torch.bincount(du.to(torch.int64), minlength=w["users"]); icnt = torch.bincount(di.to(torch.int64), minlength=w["items"])
    ucnt.cpu(); icnt.cpu()
    t2 = T()
    R = csr.build_csr(du, di, dr, w["users"], counts=ucnt)
    t3 = T()
    Rt = csr.build_csr(di, du, dr, w["items"], counts=icnt)
    t4 = T()
    pR = csr.AlsPlanHandle(R, w["rank"], n_src=w["items"]); pRt = csr.AlsPlanHandle(Rt, w["rank"], n_src=w["users"])
    t5 = T()
    print(f"h2d {1e3*(t1-t0):.2f} counts {1e3*(t2-t1):.2f} csr_user {1e3*(t3-t2):.2f} csr_item {1e3*(t4-t3):.2f} plans {1e3*(t5-t4):.2f} ms")
