"""Replay of the reference's structural pins on table naming / sharing:
tzrec/modules/embedding_test.py:774-945 `test_embedding_name_suffix` (six cases) and
`test_embedding_name_suffix_collision_raises` (:947-985), with the same features (:77-166) and the same
feature-group definitions, written in the text format this repo's loader parses."""
import pytest
import torch

from torcheasyrec_b200.config import parse_text
from torcheasyrec_b200.embedding_group import EmbeddingGroup
from torcheasyrec_b200.features import create_features

FEATURES = """
feature_configs { id_feature { feature_name: "cat_a" embedding_dim: 16 expression: "item:cat_a" num_buckets: 100 } }
feature_configs { id_feature { feature_name: "cat_b" embedding_dim: 8 expression: "item:cat_b" num_buckets: 1000 } }
feature_configs { raw_feature { feature_name: "int_a" expression: "item:int_a" } }
feature_configs { sequence_feature { sequence_name: "click_seq"
    features { id_feature { feature_name: "cat_a" expression: "item:cat_a" embedding_dim: 16 num_buckets: 100 } }
    features { id_feature { feature_name: "cat_b" expression: "item:cat_b" embedding_dim: 8 num_buckets: 1000 } }
    features { raw_feature { feature_name: "int_a" expression: "item:int_a" } } } }
feature_configs { sequence_feature { sequence_name: "buy_seq"
    features { id_feature { feature_name: "cat_a" expression: "item:cat_a" embedding_dim: 16 num_buckets: 100 } }
    features { raw_feature { feature_name: "int_a" expression: "item:int_a" } } } }
"""


def _group(name, feats, gtype, suffix=None, extra=""):
    names = " ".join(f'feature_names: "{f}"' for f in feats)
    sfx = f' embedding_name_suffix: "{suffix}"' if suffix is not None else ""
    return f'feature_groups {{ group_name: "{name}" {names} group_type: {gtype}{sfx} {extra} }}\n'


class _Enc(torch.nn.Module):            # stand-in for the sequence encoder the DEEP group needs (naming only)
    def output_dim(self):
        return 16

    def forward(self, x):
        raise NotImplementedError


def _names(groups_text):
    cfg = parse_text(FEATURES + "model_config { " + groups_text + " }")
    features = create_features(list(cfg.feature_configs))
    eg = EmbeddingGroup(features, list(cfg.model_config.feature_groups), device=torch.device("cpu"),
                        seq_encoder_factory=lambda c, dims: _Enc())
    ebc = {c.name for impl in eg.emb_impls.values() if impl.has_sparse for c in impl.ebc.embedding_bag_configs()}
    ec = {c.name for impl in eg.seq_emb_impls.values() for e in impl.ec_dict.values() for c in e.embedding_configs()}
    return ebc, ec


CASES = {
    # name: (groups, expected_ebc, forbidden_ebc, expected_ec, forbidden_ec)  — embedding_test.py:776-838
    "distinct_suffix_per_tower": (
        _group("deep_a", ["cat_a", "cat_b", "int_a"], "DEEP", "tower_a") +
        _group("deep_b", ["cat_a", "cat_b", "int_a"], "DEEP", "tower_b"),
        {"cat_a_emb_tower_a", "cat_a_emb_tower_b", "cat_b_emb_tower_a", "cat_b_emb_tower_b"}, {"cat_a_emb", "cat_b_emb"},
        set(), set()),
    "same_suffix_shares": (
        _group("deep_a", ["cat_a", "cat_b", "int_a"], "DEEP", "shared") +
        _group("deep_b", ["cat_a", "cat_b", "int_a"], "DEEP", "shared"),
        {"cat_a_emb_shared", "cat_b_emb_shared"}, {"cat_a_emb", "cat_b_emb"}, set(), set()),
    "wide_plus_suffix": (
        _group("wide", ["cat_a", "cat_b"], "WIDE", "tower_a"),
        {"cat_a_emb_wide_tower_a", "cat_b_emb_wide_tower_a"}, {"cat_a_emb_wide", "cat_a_emb"}, set(), set()),
    "nested_seq_inherits_parent": (
        _group("parent", ["cat_a"], "DEEP", "tower_a",
               'sequence_groups { group_name: "seq" feature_names: "cat_a" feature_names: "click_seq__cat_a" } '
               'sequence_encoders { simple_attention { input: "seq" } }'),
        {"cat_a_emb_tower_a"}, {"cat_a_emb"},
        {"cat_a_emb_tower_a", "click_seq__cat_a_emb_tower_a"}, {"cat_a_emb", "click_seq__cat_a_emb"}),
    "explicit_child_overrides_parent": (
        _group("parent", ["cat_a"], "DEEP", "parent",
               'sequence_groups { group_name: "seq" feature_names: "cat_a" feature_names: "click_seq__cat_a" '
               'embedding_name_suffix: "child" } sequence_encoders { simple_attention { input: "seq" } }'),
        {"cat_a_emb_parent"}, set(),
        {"cat_a_emb_child", "click_seq__cat_a_emb_child"}, {"cat_a_emb_parent", "click_seq__cat_a_emb_parent"}),
    "toplevel_sequence_with_suffix": (
        _group("toplevel_seq", ["cat_a", "cat_b", "click_seq__cat_a", "click_seq__cat_b"], "SEQUENCE", "tower_a"),
        set(), set(),
        {"cat_a_emb_tower_a", "cat_b_emb_tower_a", "click_seq__cat_a_emb_tower_a", "click_seq__cat_b_emb_tower_a"},
        {"cat_a_emb", "click_seq__cat_a_emb"}),
}


@pytest.mark.parametrize("case", list(CASES))
def test_embedding_name_suffix(case):
    groups, exp_ebc, forb_ebc, exp_ec, forb_ec = CASES[case]
    ebc, ec = _names(groups)
    assert exp_ebc <= ebc, f"missing from ebc: {exp_ebc - ebc}; got {ebc}"
    assert not (forb_ebc & ebc), f"forbidden in ebc: {forb_ebc & ebc}"
    assert exp_ec <= ec, f"missing from ec: {exp_ec - ec}; got {ec}"
    assert not (forb_ec & ec), f"forbidden in ec: {forb_ec & ec}"


def test_embedding_name_suffix_collision_raises():
    text = """
feature_configs { id_feature { feature_name: "cat_a" embedding_dim: 16 num_buckets: 100 } }
feature_configs { id_feature { feature_name: "cat_collide" embedding_dim: 16 num_buckets: 100 embedding_name: "cat_a_emb_x" } }
model_config {
  feature_groups { group_name: "deep_a" feature_names: "cat_a" group_type: DEEP embedding_name_suffix: "x" }
  feature_groups { group_name: "deep_b" feature_names: "cat_collide" group_type: DEEP }
}"""
    cfg = parse_text(text)
    features = create_features(list(cfg.feature_configs))
    with pytest.raises(ValueError, match="different embedding_name_suffix"):
        EmbeddingGroup(features, list(cfg.model_config.feature_groups), device=torch.device("cpu"))
