"""Replay of the reference's EmbeddingGroup unit tests on this repo's classes (CPU, oracle backend as compute):
tzrec/modules/embedding_test.py:209-256 `test_embedding_group_impl` and :273-425
`test_sequence_embedding_group_impl` (plain, non-ZCH, single-value parametrisation) — the same features, groups,
KJT fixtures (`values=[1..7], lengths=[1,2,1,3]`; `values=range(24), lengths=[1,1,1,1,3,3,3,3,2,2,2,2]`) and the same
assertions on group dims, output shapes and NaN-freeness."""
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_backend import OracleKernels  # noqa: E402

from torcheasyrec_b200 import functional as Fn  # noqa: E402
from torcheasyrec_b200.config import Message, parse_text  # noqa: E402
from torcheasyrec_b200.embedding_group import EmbeddingGroupImpl, SequenceEmbeddingGroupImpl  # noqa: E402
from torcheasyrec_b200.features import create_features  # noqa: E402
from torcheasyrec_b200.sparse import KeyedJaggedTensor, KeyedTensor  # noqa: E402
from test_embedding_naming import FEATURES as SEQ_FEATURES  # noqa: E402  (embedding_test.py:77-166)

PLAIN_FEATURES = """
feature_configs { id_feature { feature_name: "cat_a" embedding_dim: 16 num_buckets: 100 } }
feature_configs { id_feature { feature_name: "cat_b" embedding_dim: 8 num_buckets: 1000 } }
feature_configs { id_feature { feature_name: "cat_c" embedding_dim: 12 num_buckets: 1000 } }
feature_configs { raw_feature { feature_name: "int_a" } }
"""


def _fg(name, feats, gtype):
    names = " ".join(f'feature_names: "{f}"' for f in feats)
    return f'feature_groups {{ group_name: "{name}" {names} group_type: {gtype} }}\n'


def _seq_group(name, feats):
    m = Message("SeqGroupConfig")
    m.group_name = name
    for f in feats:
        m.feature_names.append(f)
    return m


def test_embedding_group_impl():
    cfg = parse_text(PLAIN_FEATURES + "model_config { " + _fg("wide", ["cat_a", "cat_b"], "WIDE") +
                     _fg("deep", ["cat_a", "cat_b", "int_a"], "DEEP") + " }")
    features = create_features(list(cfg.feature_configs))
    eg = EmbeddingGroupImpl(features, list(cfg.model_config.feature_groups), device=torch.device("cpu"))
    assert eg.group_dims("wide") == [4, 4]
    assert eg.group_dims("deep") == [16, 8, 1]
    assert eg.group_total_dim("wide") == 8
    assert eg.group_total_dim("deep") == 25
    assert eg.group_feature_dims("wide") == OrderedDict({"cat_a": 4, "cat_b": 4})
    assert eg.group_feature_dims("deep") == OrderedDict({"cat_a": 16, "cat_b": 8, "int_a": 1})
    sparse = KeyedJaggedTensor.from_lengths_sync(keys=["cat_a", "cat_b"], values=torch.tensor([1, 2, 3, 4, 5, 6, 7]),
                                                 lengths=torch.tensor([1, 2, 1, 3], dtype=torch.int32))
    dense = KeyedTensor.from_tensor_list(keys=["int_a"], tensors=[torch.tensor([[0.2], [0.3]])])
    with Fn.use_backend(OracleKernels()), torch.no_grad():
        result = eg(sparse, dense)
    assert result["wide"].size() == (2, 8)
    assert result["deep"].size() == (2, 25)
    # the raw feature rides through the regroup untouched, behind the two embeddings
    assert torch.equal(result["deep"][:, 24], torch.tensor([0.2, 0.3]))


def test_sequence_embedding_group_impl():
    cfg = parse_text(SEQ_FEATURES + "model_config { " +
                     _fg("click", ["cat_a", "cat_b", "int_a", "click_seq__cat_a", "click_seq__cat_b", "click_seq__int_a"],
                         "SEQUENCE") +
                     _fg("buy", ["cat_a", "int_a", "buy_seq__cat_a", "buy_seq__int_a"], "SEQUENCE") + " }")
    features = create_features(list(cfg.feature_configs))
    groups = list(cfg.model_config.feature_groups) + [
        _seq_group("deep___click_all", ["cat_a", "cat_b", "int_a", "click_seq__cat_a", "click_seq__cat_b", "click_seq__int_a"]),
        _seq_group("deep___click_other", ["cat_a", "int_a", "click_seq__cat_a", "click_seq__int_a"]),
        _seq_group("deep___click_no_query", ["click_seq__cat_a", "click_seq__int_a"]),
    ]
    eg = SequenceEmbeddingGroupImpl(features, groups, device=torch.device("cpu"))
    expect_dims = {"click.sequence": [16, 8, 1], "click.query": [16, 8, 1], "buy.sequence": [16, 1], "buy.query": [16, 1],
                   "deep___click_all.sequence": [16, 8, 1], "deep___click_all.query": [16, 8, 1],
                   "deep___click_other.sequence": [16, 1], "deep___click_other.query": [16, 1],
                   "deep___click_no_query.sequence": [16, 1], "deep___click_no_query.query": []}
    for g, dims in expect_dims.items():
        assert eg.group_dims(g) == dims, g
        assert eg.group_total_dim(g) == sum(dims), g
    sparse = KeyedJaggedTensor.from_lengths_sync(
        keys=["cat_a", "cat_b", "click_seq__cat_a", "click_seq__cat_b", "buy_seq__cat_a", "buy_seq__cat_b"],
        values=torch.tensor(list(range(24))), lengths=torch.tensor([1, 1, 1, 1, 3, 3, 3, 3, 2, 2, 2, 2], dtype=torch.int32))
    dense = KeyedTensor.from_tensor_list(keys=["int_a"], tensors=[torch.tensor([[0.2], [0.3]])])
    seq_dense = KeyedJaggedTensor.from_lengths_sync(
        keys=["click_seq__int_a", "buy_seq__int_a"], values=torch.tensor([[x] for x in range(10)], dtype=torch.float32),
        lengths=torch.tensor([3, 3, 2, 2], dtype=torch.int32)).to_dict()
    with Fn.use_backend(OracleKernels()), torch.no_grad():
        result = eg(sparse, dense, seq_dense)
    shapes = {"click.query": (2, 25), "click.sequence": (2, 3, 25), "click.sequence_length": (2,),
              "buy.query": (2, 17), "buy.sequence": (2, 2, 17), "buy.sequence_length": (2,),
              "deep___click_all.query": (2, 25), "deep___click_all.sequence": (2, 3, 25),
              "deep___click_all.sequence_length": (2,), "deep___click_other.query": (2, 17),
              "deep___click_other.sequence": (2, 3, 17), "deep___click_other.sequence_length": (2,),
              "deep___click_no_query.sequence": (2, 3, 17), "deep___click_no_query.sequence_length": (2,)}
    for k, shp in shapes.items():
        assert tuple(result[k].size()) == shp, (k, tuple(result[k].size()))
        assert not torch.any(torch.isnan(result[k].float())).item(), k
    assert "deep___click_no_query.query" not in result


import pytest  # noqa: E402


@pytest.mark.parametrize("pooling", ["sum", "mean"])
def test_sequence_embedding_group_impl_mulval(pooling):
    """embedding_test.py:273-425 with has_mulval=True: cat_a carries several ids per value (value_dim 0) both as a
    plain feature and inside the sequences; per-step pooling through sequence_mulval_lengths
    (`values=[1,0,3,1,2,2,1,2,2,2], lengths=[3,3,2,2]`), ids `range(30)`."""
    feats = SEQ_FEATURES.replace('feature_name: "cat_a" embedding_dim: 16 expression: "item:cat_a" num_buckets: 100',
                                 f'feature_name: "cat_a" embedding_dim: 16 expression: "item:cat_a" num_buckets: 100 '
                                 f'value_dim: 0 pooling: "{pooling}"')
    feats = feats.replace('feature_name: "cat_a" expression: "item:cat_a" embedding_dim: 16 num_buckets: 100',
                          f'feature_name: "cat_a" expression: "item:cat_a" embedding_dim: 16 num_buckets: 100 '
                          f'value_dim: 0 pooling: "{pooling}"')
    assert feats.count("value_dim: 0") == 3
    cfg = parse_text(feats + "model_config { " +
                     _fg("click", ["cat_a", "cat_b", "int_a", "click_seq__cat_a", "click_seq__cat_b", "click_seq__int_a"],
                         "SEQUENCE") +
                     _fg("buy", ["cat_a", "int_a", "buy_seq__cat_a", "buy_seq__int_a"], "SEQUENCE") + " }")
    features = create_features(list(cfg.feature_configs))
    eg = SequenceEmbeddingGroupImpl(features, list(cfg.model_config.feature_groups), device=torch.device("cpu"))
    assert eg.has_mulval_seq
    sparse = KeyedJaggedTensor.from_lengths_sync(
        keys=["cat_a", "cat_b", "click_seq__cat_a", "click_seq__cat_b", "buy_seq__cat_a", "buy_seq__cat_b"],
        values=torch.tensor(list(range(30))),
        lengths=torch.tensor([2, 0, 1, 1, 4, 5, 3, 3, 3, 4, 2, 2], dtype=torch.int32))
    mulval = KeyedJaggedTensor.from_lengths_sync(keys=["click_seq__cat_a", "buy_seq__cat_a"],
                                                 values=torch.tensor([1, 0, 3, 1, 2, 2, 1, 2, 2, 2]),
                                                 lengths=torch.tensor([3, 3, 2, 2], dtype=torch.int32))
    dense = KeyedTensor.from_tensor_list(keys=["int_a"], tensors=[torch.tensor([[0.2], [0.3]])])
    seq_dense = KeyedJaggedTensor.from_lengths_sync(
        keys=["click_seq__int_a", "buy_seq__int_a"], values=torch.tensor([[x] for x in range(10)], dtype=torch.float32),
        lengths=torch.tensor([3, 3, 2, 2], dtype=torch.int32)).to_dict()
    with Fn.use_backend(OracleKernels()), torch.no_grad():
        result = eg(sparse, dense, seq_dense, mulval)
    for k, shp in {"click.query": (2, 25), "click.sequence": (2, 3, 25), "click.sequence_length": (2,),
                   "buy.query": (2, 17), "buy.sequence": (2, 2, 17), "buy.sequence_length": (2,)}.items():
        assert tuple(result[k].size()) == shp, (k, tuple(result[k].size()))
        assert not torch.any(torch.isnan(result[k].float())).item(), k
    # values: click_seq__cat_a of sample 0 holds ids 4..7 in steps of [1, 0, 3] ids: step 0 = row 4, step 1 has no id
    # -> zeros, step 2 pools rows 5, 6, 7
    ec = eg.ec_dict["16"]
    t = [c.name for c in ec.embedding_configs()].index("click_seq__cat_a_emb")
    w = ec.table_weight(t)
    got = result["click.sequence"][0, :, :16]
    assert torch.count_nonzero(got[1]) == 0
    want = w[[5, 6, 7]].sum(0) if pooling == "sum" else w[[5, 6, 7]].mean(0)
    torch.testing.assert_close(got[2], want, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(got[0], w[4], rtol=1e-6, atol=1e-7)


def test_batch_flattening_example_of_the_reference():
    """tzrec/datasets/utils.py:311-321: the multi-value sequence `click_seq` = [[[3, 4], [5]], [6, [7, 8]]] travels as
    values [3..8], accumulated lengths [3, 3], key_lengths [2, 1, 1, 2], seq_lengths [2, 2]; the sequence output must
    be the per-step sums of the table rows."""
    cfg = parse_text("""
feature_configs { sequence_feature { sequence_name: "click_seq"
    features { id_feature { feature_name: "item" embedding_dim: 4 num_buckets: 10 value_dim: 0 pooling: "sum" } } } }
model_config { feature_groups { group_name: "g" feature_names: "click_seq__item" group_type: SEQUENCE } }""")
    features = create_features(list(cfg.feature_configs))
    eg = SequenceEmbeddingGroupImpl(features, list(cfg.model_config.feature_groups), device=torch.device("cpu"))
    sparse = KeyedJaggedTensor.from_lengths_sync(keys=["click_seq__item"], values=torch.tensor([3, 4, 5, 6, 7, 8]),
                                                 lengths=torch.tensor([3, 3], dtype=torch.int32))
    mulval = KeyedJaggedTensor.from_lengths_sync(keys=["click_seq__item"], values=torch.tensor([2, 1, 1, 2]),
                                                 lengths=torch.tensor([2, 2], dtype=torch.int32))
    with Fn.use_backend(OracleKernels()), torch.no_grad():
        res = eg(sparse, None, {}, mulval)
    w = eg.ec_dict["4"].table_weight(0)
    assert res["g.sequence"].shape == (2, 2, 4) and res["g.sequence_length"].tolist() == [2, 2]
    torch.testing.assert_close(res["g.sequence"][0, 0], w[3] + w[4])
    torch.testing.assert_close(res["g.sequence"][0, 1], w[5])
    torch.testing.assert_close(res["g.sequence"][1, 0], w[6])
    torch.testing.assert_close(res["g.sequence"][1, 1], w[7] + w[8])
