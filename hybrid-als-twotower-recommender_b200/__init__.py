"""hybrid-als-twotower-recommender_b200: B200-native ALS half-step + hybrid top-k scoring
behind the reference's Python API (src/__init__.py:40-63 export names).

Import name: `hybrid_als_twotower_recommender_b200` (the shim module at the repo root loads
this directory, whose name contains hyphens, under that name)."""
__version__ = "0.1.0"

from .als_model import ALSModel
from .hybrid_system import HybridRecommendationSystem
from .two_tower_model import TwoTowerModel

DEFAULT_CONFIG = {   # src/__init__.py:94-109
    "ALS_PARAMS": {"rank": 10, "max_iter": 10, "reg_param": 0.1, "cold_start_strategy": "drop"},
    "TWO_TOWER_PARAMS": {"embedding_size": 50, "learning_rate": 0.001},
    "EVALUATION_PARAMS": {"k_values": [5, 10, 15, 20], "top_k": 5},
}


def get_default_config():
    return DEFAULT_CONFIG.copy()


__all__ = ["HybridRecommendationSystem", "ALSModel", "TwoTowerModel", "DEFAULT_CONFIG", "get_default_config"]
