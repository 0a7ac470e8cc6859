"""Only the one helper the hot path touches.  The reference imports `get_item_features` from
its data_preprocessing module (src/als_model.py:17) but never defines it; this is the minimal
definition consistent with how it is used (src/als_model.py:48,84,95-100): a dict
itemId -> {'features': vector, 'rating': mean rating}.  The rest of the reference's ETL
(src/data_preprocessing.py:22-134) is out of scope."""
from __future__ import annotations

import numpy as np

_FEATURE_COLUMNS = ("manufacturer_id", "category_id", "price", "average_review_rating")


def get_item_features(data):
    cols = [c for c in _FEATURE_COLUMNS if c in data.columns]
    g = data.groupby("itemId")[cols].mean()
    ratings = g["average_review_rating"].to_numpy() if "average_review_rating" in g else np.zeros(len(g))
    feats = g.to_numpy(dtype=np.float64)
    return {item: {"features": feats[n], "rating": float(ratings[n])} for n, item in enumerate(g.index)}
