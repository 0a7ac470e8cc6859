"""TwoTowerModel -- drop-in for the reference's src/two_tower_model.py (class at :17-167).

Same constructor, attributes and method names.  The forward pass that the reference runs
through Keras `Model.predict` (two_tower_model.py:145) is evaluated by the tower kernels of
libhals_b200.so: hals_tower_item / hals_tower_user precompute the tower outputs once per id
(the reference recomputes the item tower for every user) and hals_score_one_user takes the
dot products.  `self.model` is a TowerWeights holder instead of a keras.Model.

Training (two_tower_model.py:91-121) is outside the north-star path; `train` is kept so the
class stays usable end to end and runs a plain torch Adam/MSE loop over the same graph.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
from sklearn.preprocessing import MinMaxScaler

LN_EPS = 1e-3  # Keras LayerNormalization default epsilon


class TowerParams:
    """Weights of the graph built at two_tower_model.py:38-89, as fp32 device tensors."""
    NAMES = ("user_emb", "item_emb", "manu_emb", "cat_emb", "num_w", "num_b", "out_w", "out_b",
             "user_ln_g", "user_ln_b", "item_ln_g", "item_ln_b")

    def __init__(self, tensors: dict, embedding_size: int):
        self.t = tensors
        self.embedding_size = int(embedding_size)

    @classmethod
    def keras_init(cls, num_users, num_items, num_manufacturers, num_categories, embedding_size, device, seed=0):
        """Keras default initialisers: Embedding U(-0.05,0.05); Dense Glorot-uniform, zero bias;
        LayerNormalization gamma=1, beta=0."""
        import torch
        g = torch.Generator(device="cpu").manual_seed(seed)
        E = embedding_size

        def emb(n, d):
            return (torch.rand((n, d), generator=g) * 0.1 - 0.05)

        def glorot(i, o):
            lim = (6.0 / (i + o)) ** 0.5
            return (torch.rand((i, o), generator=g) * 2 - 1) * lim

        t = {
            "user_emb": emb(num_users, E), "item_emb": emb(num_items, E),
            "manu_emb": emb(num_manufacturers, 8), "cat_emb": emb(num_categories, 8),
            "num_w": glorot(2, 16), "num_b": torch.zeros(16),
            "out_w": glorot(E + 8 + 8 + 16, E), "out_b": torch.zeros(E),
            "user_ln_g": torch.ones(E), "user_ln_b": torch.zeros(E),
            "item_ln_g": torch.ones(E), "item_ln_b": torch.zeros(E),
        }
        return cls({k: v.float().contiguous().to(device) for k, v in t.items()}, E)

    @classmethod
    def from_numpy(cls, arrays: dict, device):
        import torch
        t = {k: torch.from_numpy(np.ascontiguousarray(arrays[k], dtype=np.float32)).to(device) for k in cls.NAMES}
        return cls(t, t["user_emb"].shape[1])

    def to_numpy(self):
        return {k: v.detach().cpu().numpy() for k, v in self.t.items()}

    def struct(self, scaler):
        from . import _native as nat
        w = nat.TowerWeights()
        w.embedding_size = self.embedding_size
        w.manu_dim = int(self.t["manu_emb"].shape[1])
        w.cat_dim = int(self.t["cat_emb"].shape[1])
        w.num_hidden = int(self.t["num_b"].shape[0])
        for k in self.NAMES:
            setattr(w, k, self.t[k].data_ptr())
        w.ln_eps = LN_EPS
        if scaler is not None and hasattr(scaler, "scale_"):
            sc, mn = scaler.scale_, scaler.min_
        else:
            sc, mn = (1.0, 1.0), (0.0, 0.0)
        w.num_scale[0], w.num_scale[1] = float(sc[0]), float(sc[1])
        w.num_offset[0], w.num_offset[1] = float(mn[0]), float(mn[1])
        # table sizes: ids outside a table get a zero embedding in the kernels instead of an out-of-bounds read
        w.num_users, w.num_items = int(self.t["user_emb"].shape[0]), int(self.t["item_emb"].shape[0])
        w.num_manufacturers, w.num_categories = int(self.t["manu_emb"].shape[0]), int(self.t["cat_emb"].shape[0])
        return w


class TwoTowerModel:
    """
    Two-Tower Model Architecture (reference docstring, two_tower_model.py:18-23)
    - User Tower: userId embedding
    - Item Tower: itemId + manufacturer + category + numeric features
    - Dot product similarity with layer normalization
    """

    def __init__(self, num_users, num_items, num_manufacturers, num_categories,
                 embedding_size=50, learning_rate=0.001):
        self.num_users = num_users
        self.num_items = num_items
        self.num_manufacturers = num_manufacturers
        self.num_categories = num_categories
        self.embedding_size = embedding_size
        self.learning_rate = learning_rate
        self.model = None
        self.scaler = MinMaxScaler()
        self.is_trained = False

    # -- graph --------------------------------------------------------------------------------
    def build_model(self, seed=0):
        import torch
        from . import _native as nat
        nat.lib()
        nat.require_cuda()
        self.model = TowerParams.keras_init(self.num_users, self.num_items, self.num_manufacturers,
                                            self.num_categories, self.embedding_size, torch.device("cuda"), seed)
        return self.model

    def _prepare_features(self, data):
        """Feature preprocessing (two_tower_model.py:123-134); the scaler is re-fitted here, as in
        the reference (`fit_transform`, :133)."""
        if data is None:
            return None
        return {
            "user_in": data["userId"].values,
            "item_id_in": data["itemId"].values,
            "manufacturer_in": data["manufacturer_id"].values,
            "category_in": data["category_id"].values,
            "numeric_in": self.scaler.fit_transform(data[["price", "average_review_rating"]]),
        }

    # -- forward (the hot path) -----------------------------------------------------------------
    def item_vectors(self, item_features, out=None):
        """Item tower outputs [n, E] (device) for the rows of a DataFrame with the columns
        itemId, manufacturer_id, category_id, price, average_review_rating."""
        import torch
        from . import _native as nat
        dev = self.model.t["item_emb"].device
        n = len(item_features)

        def col(name):
            return torch.from_numpy(np.ascontiguousarray(item_features[name].values, dtype=np.int32)).to(dev)

        ids, manu, cat = col("itemId"), col("manufacturer_id"), col("category_id")
        num = torch.from_numpy(np.ascontiguousarray(
            item_features[["price", "average_review_rating"]].values, dtype=np.float32)).to(dev)
        if out is None:
            out = torch.empty((n, self.embedding_size), dtype=torch.float32, device=dev)
        w = self.model.struct(self.scaler)
        nat.check(nat.lib().hals_tower_item(w, nat.ptr(ids), nat.ptr(manu), nat.ptr(cat), nat.ptr(num), n,
                                            nat.ptr(out), out.stride(0), nat.current_stream()), "hals_tower_item")
        return out

    def user_vectors(self, user_ids, out=None):
        import torch
        from . import _native as nat
        dev = self.model.t["user_emb"].device
        ids = torch.from_numpy(np.ascontiguousarray(np.asarray(user_ids), dtype=np.int32)).to(dev)
        if out is None:
            out = torch.empty((ids.numel(), self.embedding_size), dtype=torch.float32, device=dev)
        w = self.model.struct(self.scaler)
        nat.check(nat.lib().hals_tower_user(w, nat.ptr(ids), ids.numel(), nat.ptr(out), out.stride(0),
                                            nat.current_stream()), "hals_tower_user")
        return out

    def predict_for_user(self, user_id, item_features):
        """Scores of one user against every row of `item_features` (two_tower_model.py:136-146)."""
        import torch
        from . import _native as nat
        iv = self.item_vectors(item_features)
        uv = self.user_vectors([user_id])
        out = torch.empty(iv.shape[0], dtype=torch.float32, device=iv.device)
        nat.check(nat.lib().hals_score_one_user(nat.ptr(uv), nat.ptr(iv), iv.stride(0), self.embedding_size, None,
                                                iv.shape[0], nat.ptr(out), nat.current_stream()),
                  "hals_score_one_user")
        predictions = out.cpu().numpy()
        return list(zip(item_features["itemId"], predictions.flatten()))

    # -- training (outside the north-star path; plain torch) ----------------------------------------
    def _torch_forward(self, p, f, dev):
        import torch
        import torch.nn.functional as F
        E = self.embedding_size

        def rows(table, ids):
            # an id outside its table reads as a zero embedding row: what the tower kernels (csrc/towers.cu) and a Keras
            # Embedding on the GPU do -- validation users never seen in training land here
            ids = torch.as_tensor(np.asarray(ids), device=dev).long()
            ok = (ids >= 0) & (ids < table.shape[0])
            return table[torch.where(ok, ids, torch.zeros_like(ids))] * ok.unsqueeze(1).to(table.dtype)

        u = F.layer_norm(rows(p["user_emb"], f["user_in"]), (E,), p["user_ln_g"], p["user_ln_b"], LN_EPS)
        num = torch.as_tensor(np.asarray(f["numeric_in"], dtype=np.float32), device=dev)
        h = torch.relu(num @ p["num_w"] + p["num_b"])
        cat = torch.cat([rows(p["item_emb"], f["item_id_in"]), rows(p["manu_emb"], f["manufacturer_in"]),
                         rows(p["cat_emb"], f["category_in"]), h], dim=1)
        i = F.layer_norm(cat @ p["out_w"] + p["out_b"], (E,), p["item_ln_g"], p["item_ln_b"], LN_EPS)
        return (u * i).sum(dim=1)

    def train(self, train_data, val_data=None, batch_size=256, epochs=10):
        """Adam(lr) on MSE with EarlyStopping(patience=3, restore_best_weights) when validation data
        is given (two_tower_model.py:91-121).  Returns a dict of per-epoch losses."""
        import torch
        if self.model is None:
            self.build_model()
        dev = self.model.t["user_emb"].device
        params = {k: v.clone().requires_grad_(True) for k, v in self.model.t.items()}
        opt = torch.optim.Adam(list(params.values()), lr=self.learning_rate, eps=1e-7)
        has_val = val_data is not None and len(val_data) > 0
        # Embedding ids are table rows (two_tower_model.py:47-62).  Keras on the CPU raises InvalidArgumentError for a
        # TRAINING id outside its table: checked here on the host.  (Validation rows may name users / items the tables
        # do not have -- the reference's own hyperparameter_tuning holds out whole users -- they read as zero rows.)
        for frame in (train_data,):
            for col, table in (("userId", "user_emb"), ("itemId", "item_emb"), ("manufacturer_id", "manu_emb"),
                               ("category_id", "cat_emb")):
                ids = frame[col].values
                n_rows = int(self.model.t[table].shape[0])
                if len(ids) and (ids.min() < 0 or ids.max() >= n_rows):
                    raise IndexError(f"{col}: ids must lie in [0, {n_rows}) (found {int(ids.min())}..{int(ids.max())})")
        tf = self._prepare_features(train_data)
        y = torch.as_tensor(train_data["average_review_rating"].values.astype(np.float32), device=dev)
        vf = self._prepare_features(val_data) if has_val else None
        vy = torch.as_tensor(val_data["average_review_rating"].values.astype(np.float32), device=dev) if has_val else None
        if has_val:  # _prepare_features re-fits the scaler (reference behaviour); keep the train fit for predict
            tf = self._prepare_features(train_data)
        n = len(train_data)
        history = {"loss": [], "val_loss": []}
        best, best_state, bad = float("inf"), None, 0
        g = torch.Generator(device="cpu").manual_seed(0)
        for _ in range(int(epochs)):
            perm = torch.randperm(n, generator=g).numpy()
            tot = 0.0
            for s in range(0, n, batch_size):
                b = perm[s:s + batch_size]
                fb = {k: np.asarray(v)[b] for k, v in tf.items()}
                loss = ((self._torch_forward(params, fb, dev) - y[torch.as_tensor(b, device=dev)]) ** 2).mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                tot += float(loss.detach()) * len(b)
            history["loss"].append(tot / max(n, 1))
            if has_val:
                with torch.no_grad():
                    vl = float(((self._torch_forward(params, vf, dev) - vy) ** 2).mean())
                history["val_loss"].append(vl)
                if vl < best:
                    best, bad = vl, 0
                    best_state = {k: v.detach().clone() for k, v in params.items()}
                else:
                    bad += 1
                    if bad >= 3:
                        break
        final = best_state if best_state is not None else {k: v.detach() for k, v in params.items()}
        self.model = TowerParams({k: v.contiguous() for k, v in final.items()}, self.embedding_size)
        self.is_trained = True
        return history

    # -- persistence ------------------------------------------------------------------------------
    def save_model(self, model_path="models/twotower.keras"):
        os.makedirs(os.path.dirname(model_path) or ".", exist_ok=True)
        with open(model_path, "wb") as f:
            np.savez(f, embedding_size=self.embedding_size, **self.model.to_numpy())
        with open(f"{model_path}_scaler.pkl", "wb") as f:
            pickle.dump(self.scaler, f)

    @classmethod
    def load_model(cls, model_path="models/twotower.keras"):
        import torch
        from . import _native as nat
        nat.lib()
        nat.require_cuda()
        try:
            z = np.load(model_path)
            missing = [k for k in TowerParams.NAMES if k not in z.files]
        except Exception as e:       # a Keras v3 archive (what the reference's save_model writes) is a zip of config + h5
            raise ValueError(f"{model_path} is not a tower checkpoint of this package (numpy .npz written by "
                             f"TwoTowerModel.save_model); a Keras archive has to be exported layer by layer, see "
                             f"INTEGRATION.md ({e})") from e
        if missing:
            raise ValueError(f"{model_path}: not a tower checkpoint of this package (missing arrays {missing[:3]}...); "
                             "Keras archives have to be exported layer by layer, see INTEGRATION.md")
        with open(f"{model_path}_scaler.pkl", "rb") as f:
            scaler = pickle.load(f)
        params = TowerParams.from_numpy({k: z[k] for k in TowerParams.NAMES}, torch.device("cuda"))
        loaded_model = cls(params.t["user_emb"].shape[0], params.t["item_emb"].shape[0],
                           params.t["manu_emb"].shape[0], params.t["cat_emb"].shape[0],
                           embedding_size=params.embedding_size)
        loaded_model.model = params
        loaded_model.scaler = scaler
        loaded_model.is_trained = True
        return loaded_model


def hyperparameter_tuning(train_data, param_grid, val_size=0.2, random_state=42):
    """F1@10-based grid search over {batch_size, epochs} (src/two_tower_model.py:169-236): hold out `val_size` of the
    users, train one TwoTowerModel(embedding_size=50, learning_rate=0.001) per grid entry on the rest, score the first 50
    held-out users against the held-out items with predict_for_user (the GPU tower kernels) and keep the entry with the
    best mean F1@10.  Returns the best entry (a copy) or None; an entry that raises is reported and skipped, like the
    reference.  One deviation: the reference passes `random_state` to np.random.choice, which has no such argument (its
    call raises TypeError before anything runs); the held-out users are drawn from np.random.RandomState(random_state)."""
    best_params = None
    best_f1 = 0.0
    train_users = train_data['userId'].unique()
    rng = np.random.RandomState(random_state)
    val_users = rng.choice(train_users, size=int(len(train_users) * val_size), replace=False)
    train_sub = train_data[~train_data['userId'].isin(val_users)]
    val_sub = train_data[train_data['userId'].isin(val_users)]
    num_users = train_sub['userId'].nunique()
    num_items = train_sub['itemId'].nunique()
    num_manufacturers = train_sub['manufacturer_id'].nunique()
    num_categories = train_sub['category_id'].nunique()
    for params in param_grid:
        print(f"\nTesting parameters: {params}")
        try:
            model = TwoTowerModel(num_users=num_users, num_items=num_items, num_manufacturers=num_manufacturers,
                                  num_categories=num_categories, embedding_size=50, learning_rate=0.001)
            model.train(train_sub, val_sub, batch_size=params['batch_size'], epochs=params['epochs'])
            f1_scores = []
            sample_users = val_sub['userId'].unique()[:50]
            items = val_sub[['itemId', 'manufacturer_id', 'category_id', 'price',
                             'average_review_rating']].drop_duplicates()
            for user_id in sample_users:
                rows = val_sub[val_sub['userId'] == user_id]
                actual = dict(zip(rows['itemId'], rows['average_review_rating']))
                preds = model.predict_for_user(user_id, items)
                f1_scores.append(compute_f1_score(actual, dict(preds), k=10))
            avg_f1 = np.mean(f1_scores)
            print(f"  Avg F1@10: {avg_f1:.4f}")
            if avg_f1 > best_f1:
                best_f1 = avg_f1
                best_params = params.copy()
                print(f"  New best F1@10: {best_f1:.4f}")
        except Exception as e:
            print(f"  Error with params {params}: {str(e)}")
            continue
    return best_params


def compute_f1_score(actual, pred, k=10):
    """two_tower_model.py:238-245 (same function as als_model.compute_f1_score)."""
    from .als_model import compute_f1_score as f
    return f(actual, pred, k)
