"""`compute_f1_score` under the module name src/hybrid_system.py:15 imports it from (the
reference's evaluation.py does not define it; the only definitions are als_model.py:171-177
and the identical two_tower_model.py:238-245).  The rest of the reference's evaluation.py
(RecommenderEvaluator, plots) is out of scope."""
from .als_model import compute_f1_score  # noqa: F401
