"""CPU ORACLE (test infrastructure, NOT product code) -- two-tower forward.

Restates the Keras graph the reference builds at src/two_tower_model.py:38-89
and evaluates at src/two_tower_model.py:136-146.  The layer arithmetic lives in
tensorflow==2.8.0 (requirements.txt:2), which is not under /root/reference and
is not installed here: PARITY UNPINNED (no TF to compare with, and the reference
has no fixtures).  Anchored on the reference's layer list and Keras defaults:

  user tower (two_tower_model.py:71-74): Embedding(num_users,E) -> Flatten -> LayerNormalization
  item tower (two_tower_model.py:41-64): Embedding(num_items,E), Embedding(num_manu,8),
      Embedding(num_cat,8), Dense(16,relu)(numeric[2]) -> Concatenate (that order, :60)
      -> Dense(E) -> LayerNormalization
  score (two_tower_model.py:80): Dot(axes=1)
  Keras LayerNormalization defaults: axis=-1, epsilon=1e-3, biased variance,
      gamma*(x-mean)/sqrt(var+eps)+beta.   Dense: y = x W + b.
  numeric features: MinMaxScaler fitted at train time (two_tower_model.py:133), only
      transform()-ed at predict time (:143):  x*scale_ + min_

All arithmetic fp32 (TF default dtype), reductions accumulated in fp64 then cast,
which is within 1 ulp-ish of any fp32 summation order.
"""
from __future__ import annotations

import numpy as np

LN_EPS = 1e-3


def layer_norm(x, gamma, beta, eps=LN_EPS):
    xd = x.astype(np.float64)
    mu = xd.mean(axis=-1, keepdims=True)
    var = ((xd - mu) ** 2).mean(axis=-1, keepdims=True)
    y = (xd - mu) / np.sqrt(var + eps) * gamma.astype(np.float64) + beta.astype(np.float64)
    return y.astype(np.float32)


def init_weights(num_users, num_items, num_manufacturers, num_categories, embedding_size=50, seed=0):
    """Random weights with the Keras default initialisers (Embedding U(-0.05,0.05),
    Dense Glorot-uniform / zero bias, LN gamma=1 beta=0), then perturbed biases/gains so
    that parity tests exercise every term."""
    rng = np.random.default_rng(seed)
    E = embedding_size

    def emb(n, d):
        return rng.uniform(-0.05, 0.05, size=(n, d)).astype(np.float32)

    def glorot(i, o):
        lim = np.sqrt(6.0 / (i + o))
        return rng.uniform(-lim, lim, size=(i, o)).astype(np.float32)

    cat = E + 8 + 8 + 16
    return {
        "user_emb": emb(num_users, E),
        "item_emb": emb(num_items, E),
        "manu_emb": emb(num_manufacturers, 8),
        "cat_emb": emb(num_categories, 8),
        "num_w": glorot(2, 16),
        "num_b": (0.1 * rng.standard_normal(16)).astype(np.float32),
        "out_w": glorot(cat, E),
        "out_b": (0.1 * rng.standard_normal(E)).astype(np.float32),
        "user_ln_g": (1.0 + 0.1 * rng.standard_normal(E)).astype(np.float32),
        "user_ln_b": (0.1 * rng.standard_normal(E)).astype(np.float32),
        "item_ln_g": (1.0 + 0.1 * rng.standard_normal(E)).astype(np.float32),
        "item_ln_b": (0.1 * rng.standard_normal(E)).astype(np.float32),
    }


def scale_numeric(numeric, scale, offset):
    """MinMaxScaler.transform: X*scale_ + min_ (fp64 in sklearn, cast to fp32 by Keras)."""
    return (np.asarray(numeric, dtype=np.float64) * np.asarray(scale, dtype=np.float64)
            + np.asarray(offset, dtype=np.float64)).astype(np.float32)


def user_tower(w, user_ids):
    return layer_norm(w["user_emb"][np.asarray(user_ids)], w["user_ln_g"], w["user_ln_b"])


def item_tower(w, item_ids, manufacturer_ids, category_ids, numeric_scaled):
    num = np.asarray(numeric_scaled, dtype=np.float32)
    h = (num.astype(np.float64) @ w["num_w"].astype(np.float64) + w["num_b"]).astype(np.float32)
    h = np.maximum(h, np.float32(0))
    concat = np.concatenate([w["item_emb"][np.asarray(item_ids)], w["manu_emb"][np.asarray(manufacturer_ids)],
                             w["cat_emb"][np.asarray(category_ids)], h], axis=1)
    z = (concat.astype(np.float64) @ w["out_w"].astype(np.float64) + w["out_b"]).astype(np.float32)
    return layer_norm(z, w["item_ln_g"], w["item_ln_b"])


def score(user_vecs, item_vecs):
    """Dot(axes=1) for one user against many items -> [n_items] fp32."""
    return (item_vecs.astype(np.float64) @ user_vecs.astype(np.float64).reshape(-1)).astype(np.float32)
