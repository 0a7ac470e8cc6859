"""CPU, world_size 2, gloo: the multi-GPU host logic (row sharding, factor all-gather after each
half-step, item-sharded scoring exchange) with the kernels replaced by the CPU oracle.  The sharded
result must equal the single-process oracle result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import als_oracle, c_oracle, hybrid_oracle


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _oracle_half_step(shard, plan, src, dst_full, k, reg, implicit, alpha, gram):
    out = c_oracle.als_half_step(shard.rowptr_host, shard.colidx.numpy(), shard.vals.numpy(), src.numpy(), reg,
                                 implicit, alpha)
    dst_full[shard.row_begin: shard.row_end] = torch.from_numpy(out)


def _cpu_merge(part_idx, part_score):
    P, U, k = part_idx.shape
    idx = part_idx.permute(1, 0, 2).reshape(U, P * k).numpy()
    sc = part_score.permute(1, 0, 2).reshape(U, P * k).numpy().astype(np.float64)
    key = np.where(idx >= 0, sc, -np.inf)
    order = np.lexsort((np.where(idx >= 0, idx, 2 ** 31 - 1), -key), axis=1)[:, :k]
    return torch.from_numpy(np.take_along_axis(idx, order, 1)), torch.from_numpy(np.take_along_axis(sc, order, 1).astype(np.float32))


def _worker(rank, world, port, case, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hybrid_als_twotower_recommender_b200  # noqa: F401
    from hybrid_als_twotower_recommender_b200 import als_engine, scoring
    z = np.load(case)
    U, I, k = int(z["U"]), int(z["I"]), int(z["k"])
    uu, ii, rr = z["u"], z["i"], z["r"]
    part = bool(z["partitioned"])
    if part:      # every rank holds only its contiguous slice of the triples; the engine routes them to the row owners
        c0, c1 = len(uu) * rank // world, len(uu) * (rank + 1) // world
        uu, ii, rr = uu[c0:c1], ii[c0:c1], rr[c0:c1]
    eng = als_engine.AlsEngine(uu, ii, rr, U, I, k, 0.1, implicit=bool(z["implicit"]), alpha=4.0,
                               device="cpu", dist_rank=rank, world=world, half_step=_oracle_half_step, make_plans=False,
                               partitioned=part)
    assert eng.nnz_total == len(z["u"])
    assert eng.R.row_begin == eng.user_bounds[rank] and eng.Rt.row_end == eng.item_bounds[rank + 1]
    eng.set_user_factors(z["X0"])
    X, Y = eng.fit(3)
    # item-sharded scoring exchange on top of the fitted factors
    ib, ie = scoring.shard_items(I, rank, world)
    Ut, It = torch.from_numpy(z["Ut"]), torch.from_numpy(z["It"])
    Sa = (X.double() @ Y[ib:ie].double().T); St = (Ut.double() @ It[ib:ie].double().T)
    ex = torch.stack([Sa.min(1).values, Sa.max(1).values, St.min(1).values, St.max(1).values], 1).float()
    ex = scoring.reduce_extrema(ex, world)
    exd = ex.double()
    def mm(S, lo, hi):
        rg = (hi - lo); sc = torch.where(rg != 0, 1.0 / torch.where(rg != 0, rg, torch.ones_like(rg)), torch.ones_like(rg))
        return (S - lo[:, None]) * sc[:, None]
    B = 0.8 * mm(Sa, exd[:, 0], exd[:, 1]) + 0.2 * mm(St, exd[:, 2], exd[:, 3])
    kk = 7
    order = torch.argsort(-B, dim=1, stable=True)[:, :kk]
    idx = (order + ib).to(torch.int32); sc = torch.gather(B, 1, order).float()
    fi, fs = scoring.exchange_topk(idx, sc, rank, world, _cpu_merge)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), X=X.numpy(), Y=Y.numpy(), fi=fi.numpy(), fs=fs.numpy(),
             ub=eng.user_bounds, ibn=eng.item_bounds)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("implicit,partitioned", [(False, False), (True, False), (False, True)])
def test_two_rank_sharded_fit_and_scoring_equal_single_process(tmp_path, implicit, partitioned):
    rng = np.random.default_rng(5)
    U, I, k, nnz = 91, 57, 6, 1400
    p = 1.0 / np.arange(1, I + 1); p /= p.sum()
    u, i = rng.integers(0, U, nnz), rng.choice(I, nnz, p=p)
    r = rng.integers(1, 6, nnz).astype(np.float32)
    X0 = als_oracle.init_factors(U, k, 3)
    Ut, It = rng.normal(size=(U, 5)).astype(np.float32), rng.normal(size=(I, 5)).astype(np.float32)
    case = str(tmp_path / "case.npz")
    np.savez(case, u=u, i=i, r=r, U=U, I=I, k=k, X0=X0, Ut=Ut, It=It, implicit=implicit, partitioned=partitioned)
    mp.spawn(_worker, args=(2, _free_port(), case, str(tmp_path)), nprocs=2, join=True)
    Xo, Yo = als_oracle.als_fit(u, i, r, U, I, k, 3, 0.1, X0, implicit=implicit, alpha=4.0,
                                half_step=c_oracle.als_half_step)
    outs = [np.load(tmp_path / f"rank{q}.npz") for q in range(2)]
    for o in outs:       # every rank holds the full, identical factor matrices after the all-gathers
        assert np.array_equal(o["X"], Xo) and np.array_equal(o["Y"], Yo)
        assert o["ub"][0] == 0 and o["ub"][-1] == U and o["ibn"][-1] == I
    wi, ws = hybrid_oracle.hybrid_topk_dense(Xo, Yo, Ut, It, 0.8, 0.2, 7)
    per = (U + 1) // 2
    got_i = np.concatenate([outs[0]["fi"], outs[1]["fi"]])[:U]
    got_s = np.concatenate([outs[0]["fs"], outs[1]["fs"]])[:U]
    assert outs[0]["fi"].shape[0] == per
    assert np.allclose(got_s, ws, atol=2e-6)
    assert (got_i == wi).mean() > 0.995


def test_balanced_row_bounds_and_csr_shards():
    import hybrid_als_twotower_recommender_b200  # noqa: F401
    from hybrid_als_twotower_recommender_b200 import csr
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 40, 1000); counts[17] = 5000
    for world in (1, 2, 3, 8):
        b = csr.balanced_row_bounds(counts, world)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0)
        per = [counts[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(per) <= counts.sum() / world + counts.max()
    rows = torch.from_numpy(np.repeat(np.arange(1000), counts)); cols = torch.arange(rows.numel()) % 77
    vals = torch.rand(rows.numel())
    full = csr.build_csr(rows, cols, vals, 1000)
    assert np.array_equal(np.diff(full.rowptr_host), counts)
    b = csr.balanced_row_bounds(counts, 3)
    parts = [csr.build_csr(rows, cols, vals, 1000, int(b[r]), int(b[r + 1])) for r in range(3)]
    assert sum(p.nnz for p in parts) == full.nnz
    assert torch.equal(torch.cat([p.colidx for p in parts]), full.colidx)      # stable order kept
    assert np.array_equal(parts[1].rowptr_host, full.rowptr_host[b[1]:b[2] + 1] - full.rowptr_host[b[1]])
