"""[weight row | Adagrad accumulator row] arenas (tzk_opt_args.interleaved, kernels.build_layout(interleaved=True)):
the host logic around them — layout arithmetic, per-table views, state_dict / optimizer-state round trips — and, through
the oracle backend's strided numpy views, that a model trained on such an arena follows the dense-row model bit for bit.
(The CUDA kernels over the same layout: tests/test_kernels_gpu.py::test_interleaved_*.)"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_backend import OracleKernels  # noqa: E402

from torcheasyrec_b200 import functional as Fn  # noqa: E402
from torcheasyrec_b200.embedding_modules import (EmbeddingBagCollection, EmbeddingBagConfig, PoolingType,  # noqa: E402
                                                 SparseOptimizerSpec)
from torcheasyrec_b200.engine import Pipeline  # noqa: E402
from torcheasyrec_b200.kernels import OPT_ADAGRAD, OPT_SGD, build_layout  # noqa: E402


def test_layout_arithmetic():
    lay = build_layout([5, 3, 7], [16, 4, 16], [0, 1, 1, 2], [0, 0, 1, 0], interleaved=True)
    assert lay.interleaved and lay.stride == [32, 8, 8, 32]
    assert lay.w_off[0] == 0 and lay.w_off[1] == 160 and lay.w_off[1] == lay.w_off[2]      # 5 lines of 32 floats
    assert lay.w_off[3] % 32 == 0 and lay.w_off[3] >= 160 + 3 * 8                        # table starts on a 128-B line
    assert lay.key_base == [0, 5, 5, 8] and lay.total_keys == 15 and lay.col == [0, 16, 20, 24]
    assert lay.vec_ok == 1
    dense = build_layout([5, 3, 7], [16, 4, 16], [0, 1, 1, 2], [0, 0, 1, 0])
    assert dense.stride is None and not dense.interleaved and dense.row_stride(1) == 4


def _ebc():
    cfgs = [EmbeddingBagConfig(num_embeddings=11, embedding_dim=16, name="a", feature_names=["fa"]),
            EmbeddingBagConfig(num_embeddings=5, embedding_dim=4, name="b", feature_names=["fb", "fb2"],
                               pooling=PoolingType.MEAN)]
    return EmbeddingBagCollection(cfgs, device="cpu")


def test_collection_views_and_state_round_trip(monkeypatch):
    monkeypatch.setenv("TZK_INTERLEAVE", "force")
    torch.manual_seed(0)
    m = _ebc()
    w0 = [m.table_weight(t).clone() for t in range(2)]
    m.set_optimizer(SparseOptimizerSpec(kind=OPT_ADAGRAD, initial_accumulator_value=0.25))
    assert m.layout.interleaved and m.opt_state is None
    for t in range(2):
        assert torch.equal(m.table_weight(t), w0[t])                       # weights survive the re-layout
        assert torch.all(m.table_state(t) == 0.25) and m.table_state(t).shape == w0[t].shape
    assert torch.equal(m.dense_weights(), torch.cat([w.reshape(-1) for w in w0]))
    # state_dict: per-table keys, dense shapes; loads into a dense-row twin and back
    sd = m.state_dict()
    assert set(sd) == {"embedding_bags.a.weight", "embedding_bags.b.weight"} and sd["embedding_bags.a.weight"].shape == (11, 16)
    monkeypatch.setenv("TZK_INTERLEAVE", "0")
    d = _ebc()
    d.set_optimizer(SparseOptimizerSpec(kind=OPT_ADAGRAD))
    assert not d.layout.interleaved
    d.load_state_dict(sd)
    assert torch.equal(d.weights.data, m.dense_weights())
    m.table_state(0).fill_(3.0)
    d.load_fused_optimizer_state_dict(m.fused_optimizer_state_dict())
    assert torch.all(d.table_state(0) == 3.0) and torch.all(d.table_state(1) == 0.25)
    m.load_fused_optimizer_state_dict({k: v * 2 for k, v in d.fused_optimizer_state_dict().items()})
    assert torch.all(m.table_state(0) == 6.0) and torch.equal(m.table_weight(0), w0[0])
    # a second set_optimizer with another kind goes back to dense rows and keeps the weights
    m.set_optimizer(SparseOptimizerSpec(kind=OPT_SGD))
    assert not m.layout.interleaved and torch.equal(m.weights.data[:11 * 16].view(11, 16), w0[0])


@pytest.mark.parametrize("name", ["dlrm_criteo", "deepfm_criteo", "multi_tower_din_taobao"])
def test_model_on_interleaved_arena_follows_dense_rows(name, monkeypatch):
    monkeypatch.setenv("TZK_INTERLEAVE", "0")
    a = Pipeline(name, device="cpu", max_rows=300, seed=4)
    monkeypatch.setenv("TZK_INTERLEAVE", "force")
    b = Pipeline(name, device="cpu", max_rows=300, seed=4)
    colls_a, colls_b = a.model.sparse_collections(), b.model.sparse_collections()
    assert all(c.layout.interleaved for c in colls_b) and not any(c.layout.interleaved for c in colls_a)
    b.model.load_state_dict(a.model.state_dict())
    with Fn.use_backend(OracleKernels()):
        for it in range(3):
            batch = a.synthetic_batch(64, seed=it)
            la, lb = float(a.eager_step(batch)), float(b.eager_step(batch))
            assert la == lb
    for ca, cb in zip(colls_a, colls_b):
        for t in range(len(ca._configs)):
            if t in ca._table_off:
                assert np.array_equal(ca.table_weight(t).numpy(), cb.table_weight(t).numpy())
                assert np.array_equal(ca.table_state(t).numpy(), cb.table_state(t).numpy())


def test_dcp_checkpoint_round_trip_between_interleaved_and_dense(tmp_path, monkeypatch):
    """checkpoint.save_model / restore_model see per-table strided views of an interleaved arena: a checkpoint written
    from it restores into an interleaved AND into a dense-row model (weights and Adagrad accumulators)."""
    from torcheasyrec_b200 import checkpoint

    monkeypatch.setenv("TZK_INTERLEAVE", "force")
    a = Pipeline("dlrm_criteo", device="cpu", max_rows=200, seed=1)
    with Fn.use_backend(OracleKernels()):
        a.eager_step(a.synthetic_batch(64, seed=0))
    checkpoint.save_model(str(tmp_path / "ck"), a.model, a.dense_optimizer)
    b = Pipeline("dlrm_criteo", device="cpu", max_rows=200, seed=9)
    monkeypatch.setenv("TZK_INTERLEAVE", "0")
    c = Pipeline("dlrm_criteo", device="cpu", max_rows=200, seed=9)
    assert b.model.sparse_collections()[0].layout.interleaved and not c.model.sparse_collections()[0].layout.interleaved
    for p in (b, c):
        checkpoint.restore_model(str(tmp_path / "ck"), p.model, p.dense_optimizer)
        ca, cp = a.model.sparse_collections()[0], p.model.sparse_collections()[0]
        for t in range(len(ca._configs)):
            assert torch.equal(ca.table_weight(t), cp.table_weight(t))
            assert torch.equal(ca.table_state(t), cp.table_state(t))
            assert float(ca.table_state(t).abs().sum()) > 0
