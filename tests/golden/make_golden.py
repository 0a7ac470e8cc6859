"""Writes the committed fixtures under tests/golden/.  Run in the BUILD container only
(`python tests/golden/make_golden.py`): the hybrid fixtures are produced by executing the
reference's own, unmodified src/hybrid_system.py from /root/reference (oracle/ref_loader.py);
the ALS / tower fixtures are minted from the restated oracle (parity unpinned for those --
see oracle/als_oracle.py, oracle/towers_oracle.py)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import als_oracle, ref_loader, towers_oracle  # noqa: E402


def hybrid_cases():
    rng = np.random.default_rng(20261018)
    cases = []

    def add(name, als, tt, top_k=5, als_f1=None, tt_f1=None, actual=None):
        items = list(range(len(als)))
        a = [(i, float(x)) for i, x in zip(items, als)]                 # Spark -> Python float
        t = [(i, np.float32(x)) for i, x in zip(items, tt)]             # Keras -> np.float32
        out, f1 = ref_loader.reference_recommend({7: a}, {7: t}, 7, items, actual_ratings=actual,
                                                 top_k=top_k, als_f1=als_f1, tt_f1=tt_f1)
        cases.append({
            "name": name, "als": [float(x) for x in als], "tt": [float(np.float32(x)) for x in tt],
            "top_k": top_k, "als_f1_in": als_f1, "tt_f1_in": tt_f1,
            "actual": None if actual is None else {str(k): float(v) for k, v in actual.items()},
            "out_items": [int(i) for i, _ in out], "out_scores": [float(s) for _, s in out],
            "f1_after": [float(f1[0]), float(f1[1])],
        })

    add("default_weights_tt_favoured", rng.normal(3, 1, 40), rng.normal(0, 1, 40))
    add("als_favoured", rng.normal(3, 1, 40), rng.normal(0, 1, 40), als_f1=0.4, tt_f1=0.1)
    add("equal_f1_is_tt_favoured", rng.normal(3, 1, 25), rng.normal(0, 1, 25), als_f1=0.3, tt_f1=0.3)
    add("constant_als_contributes_zero", np.full(30, 2.5), rng.normal(0, 1, 30))
    add("constant_both", np.full(12, 1.0), np.full(12, -3.0), top_k=4)
    add("ties_keep_candidate_order", np.repeat([1.0, 2.0, 3.0], 6), np.tile([0.5, 0.25], 9), top_k=7)
    add("top_k_larger_than_items", rng.normal(0, 1, 3), rng.normal(0, 1, 3), top_k=5)
    add("single_item", [4.0], [0.1], top_k=5)
    add("top100_of_500", rng.normal(3, 1, 500), rng.normal(0, 2, 500), top_k=100, als_f1=0.2, tt_f1=0.1)
    act = {3: 5.0, 11: 4.0, 17: 1.0}
    add("f1_selected_from_actual_ratings", rng.normal(3, 1, 40), rng.normal(0, 1, 40), actual=act)
    return cases


def als_cases():
    out = {}
    rng = np.random.default_rng(7)
    # hand-checkable 3x3, k=2
    u = np.array([0, 0, 1, 1, 2, 2, 2]); i = np.array([0, 1, 1, 2, 0, 1, 2])
    r = np.array([5, 3, 4, 1, 2, 5, 3], np.float32)
    X0 = als_oracle.init_factors(3, 2, 1)
    for imp in (0, 1):
        X, Y = als_oracle.als_fit(u, i, r, 3, 3, 2, 3, 0.1, X0, implicit=bool(imp), alpha=2.0,
                                  half_step=als_oracle.als_half_step_loops)
        out[f"tiny_imp{imp}_X"], out[f"tiny_imp{imp}_Y"] = X, Y
    out.update(tiny_u=u, tiny_i=i, tiny_r=r, tiny_X0=X0)
    # 200x150, k=10 (the reference's rank), duplicates and empty rows included
    U, I, k, nnz = 200, 150, 10, 2500
    u = rng.integers(0, U - 5, nnz); i = rng.integers(0, I - 5, nnz)
    r = rng.integers(1, 6, nnz).astype(np.float32)
    X0 = als_oracle.init_factors(U, k, 2)
    X, Y = als_oracle.als_fit(u, i, r, U, I, k, 10, 0.1, X0)
    out.update(mid_u=u, mid_i=i, mid_r=r, mid_X0=X0, mid_X=X, mid_Y=Y,
               mid_rmse=np.float64(als_oracle.rmse(X, Y, u, i, r)))
    r2 = rng.geometric(0.4, nnz).astype(np.float32) * rng.choice([1.0, 1.0, 1.0, -1.0], nnz).astype(np.float32)
    Xi, Yi = als_oracle.als_fit(u, i, r2, U, I, k, 5, 0.05, X0, implicit=True, alpha=40.0)
    out.update(mid_r_implicit=r2, mid_Xi=Xi, mid_Yi=Yi)
    return out


def tower_case():
    w = towers_oracle.init_weights(60, 80, 17, 9, 50, seed=3)
    rng = np.random.default_rng(5)
    n = 80
    ids = rng.permutation(80)[:n]
    manu = rng.integers(0, 17, n); cat = rng.integers(0, 9, n)
    raw = np.stack([rng.uniform(1, 300, n), rng.uniform(1, 5, n)], 1)
    scale = 1.0 / (raw.max(0) - raw.min(0)); offset = -raw.min(0) * scale
    raw[:5] *= 1.5  # predict-time values may leave [0,1] (scaler fitted at train time)
    iv = towers_oracle.item_tower(w, ids, manu, cat, towers_oracle.scale_numeric(raw, scale, offset))
    uv = towers_oracle.user_tower(w, np.arange(60))
    sc = towers_oracle.score(uv[11], iv)
    d = {f"w_{k}": v for k, v in w.items()}
    d.update(ids=ids, manu=manu, cat=cat, raw=raw.astype(np.float32), scale=scale, offset=offset,
             item_vecs=iv, user_vecs=uv, scores_user11=sc)
    return d


if __name__ == "__main__":
    with open(os.path.join(HERE, "hybrid_reference_cases.json"), "w") as f:
        json.dump({"generator": "reference src/hybrid_system.py via oracle/ref_loader.py",
                   "cases": hybrid_cases()}, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "als_oracle_cases.npz"), **als_cases())
    np.savez_compressed(os.path.join(HERE, "tower_oracle_case.npz"), **tower_case())
    print("fixtures written to", HERE)
