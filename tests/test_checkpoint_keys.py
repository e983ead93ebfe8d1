"""state_dict / fused-optimizer keys follow the reference's torchrec naming
(tzrec/utils/checkpoint_util_test.py:375-396: `ebc.embedding_bags.<table>.weight`,
`state.ebc.embedding_bags.<table>.weight.<table>.momentum1`), the arena never shows up."""
import torch

from torcheasyrec_b200.embedding_modules import (EmbeddingBagCollection, EmbeddingBagConfig, EmbeddingCollection,
                                                  EmbeddingConfig, SparseOptimizerSpec)
from torcheasyrec_b200.engine import Pipeline
from torcheasyrec_b200.kernels import OPT_ADAGRAD, OPT_ROWWISE_ADAGRAD


def _ebc():
    tables = [EmbeddingBagConfig(name=n, num_embeddings=r, embedding_dim=d, feature_names=[n + "_f"])
              for n, r, d in (("large_table_0", 40, 16), ("large_table_1", 30, 16), ("small_table_0", 5, 4))]
    return torch.nn.ModuleDict({"ebc": EmbeddingBagCollection(tables, device="cpu")})


def test_ebc_keys_match_reference_fixture_names():
    m = _ebc()
    assert list(m.state_dict()) == ["ebc.embedding_bags.large_table_0.weight", "ebc.embedding_bags.large_table_1.weight",
                                    "ebc.embedding_bags.small_table_0.weight"]
    assert m.state_dict()["ebc.embedding_bags.small_table_0.weight"].shape == (5, 4)
    m["ebc"].set_optimizer(SparseOptimizerSpec(kind=OPT_ADAGRAD, lr=0.1))
    st = m["ebc"].fused_optimizer_state_dict("ebc.")
    assert list(st) == [f"state.ebc.embedding_bags.{t}.weight.{t}.momentum1"
                        for t in ("large_table_0", "large_table_1", "small_table_0")]
    assert st["state.ebc.embedding_bags.large_table_1.weight.large_table_1.momentum1"].shape == (30, 16)
    m["ebc"].set_optimizer(SparseOptimizerSpec(kind=OPT_ROWWISE_ADAGRAD, lr=0.1))
    assert m["ebc"].fused_optimizer_state_dict("ebc.")[
        "state.ebc.embedding_bags.small_table_0.weight.small_table_0.momentum1"].shape == (5,)


def test_state_dict_round_trip_and_partial_restore():
    a, b = _ebc(), _ebc()
    sd = a.state_dict()
    assert not torch.equal(a["ebc"].weights, b["ebc"].weights)
    b.load_state_dict(sd)
    assert torch.equal(a["ebc"].weights, b["ebc"].weights)
    # a reference-written checkpoint holding only some tables restores just those (strict=False)
    c = _ebc()
    before = c["ebc"].table_weight(1).clone()
    res = c.load_state_dict({"ebc.embedding_bags.large_table_0.weight": sd["ebc.embedding_bags.large_table_0.weight"]},
                            strict=False)
    assert sorted(res.missing_keys) == ["ebc.embedding_bags.large_table_1.weight", "ebc.embedding_bags.small_table_0.weight"]
    assert torch.equal(c["ebc"].table_weight(0), a["ebc"].table_weight(0))
    assert torch.equal(c["ebc"].table_weight(1), before)
    # optimizer state round trip
    a["ebc"].set_optimizer(SparseOptimizerSpec(kind=OPT_ADAGRAD, lr=0.1, initial_accumulator_value=0.5))
    b["ebc"].set_optimizer(SparseOptimizerSpec(kind=OPT_ADAGRAD, lr=0.1))
    b["ebc"].load_fused_optimizer_state_dict(a["ebc"].fused_optimizer_state_dict("ebc."), "ebc.")
    assert torch.equal(a["ebc"].opt_state, b["ebc"].opt_state)


def test_ec_keys_and_model_level_names():
    ec = torch.nn.ModuleDict({"ec": EmbeddingCollection(
        [EmbeddingConfig(name="t", num_embeddings=9, embedding_dim=8, feature_names=["f"])], device="cpu")})
    assert list(ec.state_dict()) == ["ec.embeddings.t.weight"]
    p = Pipeline("multi_tower_din_taobao", device="cpu", max_rows=50, seed=1)
    keys = list(p.model.state_dict())
    assert "embedding_group.emb_impls.__BASE__.ebc.embedding_bags.user_id_emb.weight" in keys
    assert "embedding_group.seq_emb_impls.__BASE__.ec_dict.16.embeddings.click_50_seq__adgroup_id_emb.weight" in keys
    assert not any(k.endswith(".weights") for k in keys)
    q = Pipeline("multi_tower_din_taobao", device="cpu", max_rows=50, seed=2)
    q.model.load_state_dict(p.model.state_dict())
    for a, b in zip(p.model.sparse_collections(), q.model.sparse_collections()):
        assert torch.equal(a.weights, b.weights)


def test_fp16_tables_host_logic_with_the_checker_backend():
    """SURVEY §8f N4 (FP16 half), host side: `data_type: "FP16"` in a feature config reaches the collection, the arena
    and the per-table state_dict views are halfs, lookups return fp32, the fused update moves only touched rows and
    keeps fp32 optimizer state (oracle backend as compute)."""
    import os
    import sys

    import numpy as np

    sys.path.insert(0, os.path.dirname(__file__))
    from oracle_backend import OracleKernels

    from torcheasyrec_b200 import functional as Fn
    from torcheasyrec_b200.embedding_modules import DataType
    from torcheasyrec_b200.sparse import KeyedJaggedTensor

    cfgs = [EmbeddingBagConfig(name="t", num_embeddings=50, embedding_dim=8, feature_names=["f"], data_type=DataType.FP16)]
    with Fn.use_backend(OracleKernels()):
        m = EmbeddingBagCollection(cfgs, device="cpu")
        m.set_optimizer(SparseOptimizerSpec(kind=OPT_ADAGRAD, lr=0.1))
        assert m.weights.dtype == torch.float16 and m.opt_state.dtype == torch.float32
        assert m.state_dict()["embedding_bags.t.weight"].dtype == torch.float16
        kjt = KeyedJaggedTensor.from_lengths_sync(["f"], torch.tensor([1, 2, 3, 4]), torch.tensor([1, 1, 2, 0], dtype=torch.int32))
        w0 = m.table_weight(0).clone()
        out = m(kjt).values()
        assert out.dtype == torch.float32
        np.testing.assert_array_equal(out[2].detach().numpy(), (w0[3].float() + w0[4].float()).numpy())
        out.sum().backward()
        moved = (m.table_weight(0) != w0).any(dim=1).nonzero().flatten().tolist()
        assert moved == [1, 2, 3, 4]
    import pytest

    with pytest.raises(NotImplementedError):
        EmbeddingBagCollection(cfgs + [EmbeddingBagConfig(name="u", num_embeddings=5, embedding_dim=8, feature_names=["g"])],
                               device="cpu")
