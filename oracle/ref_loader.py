"""Loads the reference's OWN src/hybrid_system.py (unmodified, read in place from
/root/reference) with stub sibling modules, so that adaptive_fusion / top-k /
compute_f1_score can be executed as the ground truth for the hybrid path.

Test infrastructure only; works only where /root/reference exists (the build
container).  The GPU box has no /root/reference: there the committed fixtures under
tests/golden/ (written by tests/golden/make_golden.py through this loader) are used.

Why stubs: the package cannot be imported as shipped -- src/als_model.py:12-17 imports
pyspark and a non-existent `get_item_features`; src/two_tower_model.py:7 imports
tensorflow; src/hybrid_system.py:15 imports a non-existent `evaluation.compute_f1_score`.
The stub for the latter is the function body AST-extracted from src/als_model.py:171-177.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

REF = os.environ.get("HALS_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF, "src", "hybrid_system.py"))


def _extract_function(path, name):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def load_reference_hybrid(pkg_name="_ref_src"):
    """Returns the reference module object for src/hybrid_system.py."""
    if not available():
        raise FileNotFoundError(REF)
    if pkg_name + ".hybrid_system" in sys.modules:
        return sys.modules[pkg_name + ".hybrid_system"]
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [os.path.join(REF, "src")]
    sys.modules[pkg_name] = pkg
    f1 = _extract_function(os.path.join(REF, "src", "als_model.py"), "compute_f1_score")

    als = types.ModuleType(pkg_name + ".als_model")
    als.ALSModel = type("ALSModel", (), {})
    als.compute_f1_score = f1
    tt = types.ModuleType(pkg_name + ".two_tower_model")
    tt.TwoTowerModel = type("TwoTowerModel", (), {})
    ev = types.ModuleType(pkg_name + ".evaluation")
    ev.compute_f1_score = f1
    for m in (als, tt, ev):
        sys.modules[m.__name__] = m
    spec = importlib.util.spec_from_file_location(
        pkg_name + ".hybrid_system", os.path.join(REF, "src", "hybrid_system.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


class FakePredictor:
    """Stands in for ALSModel / TwoTowerModel: returns a fixed score list per user."""

    def __init__(self, table):
        self.table = table

    def predict_for_user(self, user_id, all_items):
        return list(self.table[user_id])

    def stop_spark(self):
        pass


def reference_recommend(als_table, tt_table, user_id, all_items, actual_ratings=None, top_k=5,
                        als_f1=None, tt_f1=None):
    """Runs the reference's get_hybrid_recommendations on fixed per-model score lists."""
    mod = load_reference_hybrid()
    hrs = mod.HybridRecommendationSystem()
    hrs.als_model = FakePredictor(als_table)
    hrs.twotower_model = FakePredictor(tt_table)
    hrs.models_loaded = True
    if als_f1 is not None:
        hrs.als_f1_score = als_f1
    if tt_f1 is not None:
        hrs.twotower_f1_score = tt_f1
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        out = hrs.get_hybrid_recommendations(user_id, all_items, actual_ratings=actual_ratings, top_k=top_k)
    return out, (hrs.als_f1_score, hrs.twotower_f1_score)
