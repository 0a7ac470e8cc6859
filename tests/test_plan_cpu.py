"""Host-side work-plan logic (no GPU): slice-length choice of AlsPlanHandle and plan consistency.

The plan stands where Spark cuts ratings into in-blocks (ALS.scala makeBlocks, reached from
src/als_model.py:62); the kernels only ever see its arrays."""
import numpy as np
import pytest
import torch

from hybrid_als_twotower_recommender_b200 import _native as nat
from hybrid_als_twotower_recommender_b200.csr import AlsPlanHandle, CsrShard


def _shard(lens):
    lens = np.asarray(lens, dtype=np.int64)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    nnz = int(rp[-1])
    return CsrShard(0, len(lens), len(lens), torch.from_numpy(rp.copy()), torch.zeros(nnz, dtype=torch.int32),
                    torch.zeros(nnz, dtype=torch.float32), rp)


def _check_plan(plan, lens):
    h, n = plan.host, plan.n_items
    covered = np.zeros(len(lens), dtype=np.int64)
    np.add.at(covered, h["item_row"][:n], h["item_len"][:n])
    assert np.array_equal(covered, np.asarray(lens)[: len(covered)] * (np.asarray(lens) > 0))   # every rating exactly once
    assert h["item_len"][:n].max() <= plan.seg_len
    sliced = h["item_slot"][:n] >= 0
    assert int(sliced.sum()) == plan.n_slots
    assert plan.n_long == len(set(h["item_row"][:n][sliced].tolist()))


def test_default_slice_length_for_light_rows():
    lens = np.full(20000, 150)                       # user-like shard: no heavy rows -> no reason to slice finer
    default = int(nat.lib().hals_als_default_seg_len(64))
    plan = AlsPlanHandle(_shard(lens), 64, device="cpu")
    assert plan.seg_len == default and plan.n_long == 0
    _check_plan(plan, lens)


def test_small_heavy_tailed_shard_gets_finer_slices():
    lens = np.concatenate([[900_000, 300_000, 120_000], np.full(3000, 200)])   # one rank's share of a Zipf head
    default = int(nat.lib().hals_als_default_seg_len(64))
    plan = AlsPlanHandle(_shard(lens), 64, device="cpu")
    assert 512 <= plan.seg_len < default and plan.seg_len % 32 == 0
    assert plan.n_long >= 3
    _check_plan(plan, lens)


def test_big_shard_keeps_the_default_and_explicit_value_wins():
    lens = np.concatenate([[1_800_000, 900_000], np.full(26000, 650)])        # ~20M ratings: already ~8 slices per CTA
    default = int(nat.lib().hals_als_default_seg_len(64))
    assert AlsPlanHandle(_shard(lens), 64, device="cpu").seg_len == default
    plan = AlsPlanHandle(_shard([5000, 10, 0, 70]), 64, seg_len=64, device="cpu")
    assert plan.seg_len == 64 and plan.n_long == 2
    _check_plan(plan, [5000, 10, 0, 70])


def test_spark_als_model_directory_is_readable(tmp_path):
    """A checkpoint written by the reference (pyspark ALSModel.save, src/als_model.py:116-127): metadata JSON +
    userFactors / itemFactors parquet (id int, features array<float>), possibly in several part files and in any row
    order.  read_spark_als_dir returns sorted ids and fp32 factor tables."""
    import json
    import pyarrow as pa
    import pyarrow.parquet as pq
    from hybrid_als_twotower_recommender_b200.als_model import read_spark_als_dir
    rng = np.random.default_rng(0)
    rank = 6
    root = tmp_path / "als"
    (root / "metadata").mkdir(parents=True)
    (root / "metadata" / "part-00000").write_text(json.dumps({"class": "org.apache.spark.ml.recommendation.ALSModel",
                                                              "sparkVersion": "3.5.1", "rank": rank, "paramMap": {}}) + "\n")
    (root / "metadata" / "_SUCCESS").write_text("")
    want = {}
    for name, ids in (("userFactors", rng.permutation(np.arange(3, 40, 3))), ("itemFactors", rng.permutation(50)[:17])):
        f = rng.standard_normal((len(ids), rank)).astype(np.float32)
        want[name] = (ids, f)
        (root / name).mkdir()
        (root / name / "_SUCCESS").write_text("")
        half = len(ids) // 2
        for part, sl in enumerate((slice(0, half), slice(half, None))):
            t = pa.table({"id": pa.array(ids[sl].astype(np.int32)),
                          "features": pa.array([row.tolist() for row in f[sl]], type=pa.list_(pa.float32()))})
            pq.write_table(t, root / name / f"part-{part:05d}-x.snappy.parquet")
    r, uid, uf, iid, itf = read_spark_als_dir(str(root))
    assert r == rank and uf.dtype == np.float32 and uf.shape == (len(uid), rank) and itf.shape == (len(iid), rank)
    for (ids, f), (gi, gf) in ((want["userFactors"], (uid, uf)), (want["itemFactors"], (iid, itf))):
        o = np.argsort(ids)
        assert np.array_equal(gi, ids[o]) and np.array_equal(gf, f[o])
    (root / "metadata" / "part-00000").write_text(json.dumps({"rank": rank + 1}) + "\n")
    with pytest.raises(ValueError):
        read_spark_als_dir(str(root))


def test_split_factor_layout_is_padded_by_owner_rank():
    """SplitFactors (als_engine.py): rank q's rows [bounds[q], bounds[q+1]) live at [q * mx, q * mx + n_q) of the split
    matrix, so that the all-gather of freshly solved rows is one in-place equal-size collective; the last row is the
    all-zero row the ragged tail of a chunk gathers."""
    from hybrid_als_twotower_recommender_b200.als_engine import SplitFactors
    sf = SplitFactors([0, 3, 7, 7, 12], 4, "cpu", k=64)
    assert sf.mx == 5 and sf.n_rows == 20 and tuple(sf.hl.shape) == (21, 128) and sf.hl.dtype == torch.bfloat16
    assert sf.pad_of_h.tolist() == [0, 1, 2, 5, 6, 7, 8, 15, 16, 17, 18, 19]
    assert all(tuple(sf.segment(q).shape) == (5, 128) for q in range(4))
    assert sf.segment(2).data_ptr() == sf.hl[10:15].data_ptr()          # views into the one buffer (in-place all-gather)
    assert float(sf.hl.abs().sum()) == 0.0
    one = SplitFactors([0, 9], 1, "cpu", k=128)
    assert one.mx == 9 and one.n_rows == 9 and tuple(one.hl.shape) == (10, 256) and one.pad_of_h.tolist() == list(range(9))
