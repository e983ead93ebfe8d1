"""Replay of the reference's model unit tests on this repo's model shells (CPU, oracle backend as compute):
tzrec/models/dlrm_test.py:38-101 (arch_with_sparse True / False) and tzrec/models/deepfm_test.py:31-98 — same
feature configs, feature groups, model configs and the same multi-hot KJT (`values=[1..7], lengths=[1,2,1,3]`) —
plus what the reference cannot assert without golden values: the predictions equal the composition of the oracle's
pooled lookup with the (reference-pinned) dense modules."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_backend import OracleKernels  # noqa: E402

from oracle import tzk_oracle as O  # noqa: E402
from torcheasyrec_b200 import functional as Fn  # noqa: E402
from torcheasyrec_b200.batch import Batch  # noqa: E402
from torcheasyrec_b200.config import parse_text  # noqa: E402
from torcheasyrec_b200.features import create_features  # noqa: E402
from torcheasyrec_b200.rank_models import create_model  # noqa: E402
from torcheasyrec_b200.sparse import KeyedJaggedTensor, KeyedTensor  # noqa: E402

BASE = "__BASE__"


def _batch():
    sparse = KeyedJaggedTensor.from_lengths_sync(keys=["cat_a", "cat_b"], values=torch.tensor([1, 2, 3, 4, 5, 6, 7]),
                                                 lengths=torch.tensor([1, 2, 1, 3], dtype=torch.int32))
    dense = KeyedTensor.from_tensor_list(keys=["int_a"], tensors=[torch.tensor([[0.2], [0.3]])])
    return Batch(dense_features={BASE: dense}, sparse_features={BASE: sparse}, labels={})


@pytest.mark.parametrize("arch_with_sparse", [False, True])
def test_dlrm(arch_with_sparse):
    cfg = parse_text("""
feature_configs { id_feature { feature_name: "cat_a" embedding_dim: 8 num_buckets: 100 } }
feature_configs { id_feature { feature_name: "cat_b" embedding_dim: 8 num_buckets: 1000 } }
feature_configs { raw_feature { feature_name: "int_a" } }
model_config {
  feature_groups { group_name: "dense" feature_names: "int_a" group_type: DEEP }
  feature_groups { group_name: "sparse" feature_names: "cat_a" feature_names: "cat_b" group_type: DEEP }
  dlrm { dense_mlp { hidden_units: [2, 8] } final { hidden_units: [8, 4] } arch_with_sparse: %s }
  losses { binary_cross_entropy {} }
}""" % ("true" if arch_with_sparse else "false"))
    torch.manual_seed(0)
    features = create_features(list(cfg.feature_configs))
    model = create_model(cfg.model_config, features, ["label"], device=torch.device("cpu"))
    batch = _batch()
    with Fn.use_backend(OracleKernels()), torch.no_grad():
        pred = model.predict(batch)
    assert pred["logits"].size() == (2,)
    assert pred["probs"].size() == (2,)
    # value check: oracle pooled lookup (multi-hot SUM bags) -> bottom MLP -> dot interaction -> final MLP -> head
    ebc = model.sparse_collections()[0]
    tables = [ebc.table_weight(t).numpy() for t in range(2)]
    kjt = batch.sparse_features[BASE]
    pooled = O.pooled_lookup(tables, [0, 1], [O.POOL_SUM, O.POOL_SUM], kjt.values().numpy(),
                             O.lengths_to_offsets(kjt.lengths().numpy()), 2)
    with torch.no_grad():
        dense_feat = model.dense_mlp(batch.dense_features[BASE].values())
        feat = torch.cat([dense_feat.unsqueeze(1), torch.from_numpy(pooled).reshape(2, 2, 8)], dim=1)
        z = torch.bmm(feat, feat.transpose(1, 2))
        iu = torch.triu_indices(3, 3, 1)
        parts = [z[:, iu[0], iu[1]], dense_feat] + ([torch.from_numpy(pooled)] if arch_with_sparse else [])
        ref = model.output_mlp(model.final_mlp(torch.cat(parts, dim=-1))).squeeze(1)
    np.testing.assert_allclose(pred["logits"].numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(pred["probs"].numpy(), torch.sigmoid(ref).numpy(), rtol=1e-5, atol=1e-6)


def test_deepfm():
    cfg = parse_text("""
feature_configs { id_feature { feature_name: "cat_a" embedding_dim: 16 num_buckets: 100 } }
feature_configs { id_feature { feature_name: "cat_b" embedding_dim: 16 num_buckets: 1000 } }
feature_configs { raw_feature { feature_name: "int_a" } }
model_config {
  feature_groups { group_name: "wide" feature_names: "cat_a" feature_names: "cat_b" group_type: WIDE }
  feature_groups { group_name: "fm" feature_names: "cat_a" feature_names: "cat_b" group_type: DEEP }
  feature_groups { group_name: "deep" feature_names: "cat_a" feature_names: "cat_b" feature_names: "int_a" group_type: DEEP }
  deepfm { deep { hidden_units: [8, 4] } final { hidden_units: [2] } }
  losses { binary_cross_entropy {} }
}""")
    torch.manual_seed(0)
    features = create_features(list(cfg.feature_configs))
    model = create_model(cfg.model_config, features, ["label"], device=torch.device("cpu"))
    batch = _batch()
    with Fn.use_backend(OracleKernels()), torch.no_grad():
        pred = model.predict(batch)
        grouped = model.build_input(batch)
    assert pred["logits"].size() == (2,)
    assert pred["probs"].size() == (2,)
    # table set and order of the collection: the `wide` group is listed first -> `_wide` tables first
    names = [c.name for c in model.sparse_collections()[0].embedding_bag_configs()]
    assert names == ["cat_a_emb_wide", "cat_b_emb_wide", "cat_a_emb", "cat_b_emb"]
    assert grouped["wide"].shape == (2, 8) and grouped["fm"].shape == (2, 32) and grouped["deep"].shape == (2, 33)
    with torch.no_grad():     # deepfm.py:84-104 restated on the grouped tensors
        y_wide = grouped["wide"].sum(dim=1, keepdim=True)
        y_deep = model.deep_mlp(grouped["deep"])
        y_fm = torch.from_numpy(O.fm(grouped["fm"].numpy().reshape(2, 2, 16)))
        ref = model.output_mlp(model.final_mlp(torch.cat([y_wide, y_fm, y_deep], dim=1))).squeeze(1)
    np.testing.assert_allclose(pred["logits"].numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_mmoe():
    """tzrec/models/mmoe_test.py:38-119: two feature groups feeding the experts, gate MLP, two task towers."""
    cfg = parse_text("""
feature_configs { id_feature { feature_name: "cat_a" embedding_dim: 16 num_buckets: 100 } }
feature_configs { id_feature { feature_name: "cat_b" embedding_dim: 8 num_buckets: 1000 } }
feature_configs { raw_feature { feature_name: "int_a" } }
model_config {
  feature_groups { group_name: "t1" feature_names: "cat_a" feature_names: "cat_b" group_type: DEEP }
  feature_groups { group_name: "t2" feature_names: "cat_a" feature_names: "int_a" group_type: DEEP }
  mmoe {
    expert_mlp { hidden_units: [16, 8] }
    num_expert: 3
    gate_mlp { hidden_units: [4] }
    task_towers { tower_name: "t1" label_name: "label1" mlp { hidden_units: [8, 4] } losses { binary_cross_entropy {} } }
    task_towers { tower_name: "t2" label_name: "label2" mlp { hidden_units: [12, 6] } losses { binary_cross_entropy {} } }
  }
}""")
    torch.manual_seed(0)
    features = create_features(list(cfg.feature_configs))
    model = create_model(cfg.model_config, features, ["label"], device=torch.device("cpu"))
    with Fn.use_backend(OracleKernels()), torch.no_grad():
        pred = model.predict(_batch())
    for k in ("logits_t1", "probs_t1", "logits_t2", "probs_t2"):
        assert pred[k].size() == (2,), k
    assert torch.all((pred["probs_t1"] > 0) & (pred["probs_t1"] < 1))


def test_din_model_jagged_attention_equals_the_padded_reference_form(monkeypatch):
    """SURVEY §8f N3 at the model level: MultiTowerDIN with the sequence rows kept jagged (default) against the same
    model run the reference's way (TZK_DIN_JAGGED=0: longest-length read, padded [B, T, D], masked softmax) — same
    logits, same loss, same tables and dense weights after a train step (oracle backend on CPU)."""
    import sys

    import numpy as np
    import torch

    sys.path.insert(0, os.path.dirname(__file__))
    from oracle_backend import OracleKernels

    from torcheasyrec_b200 import functional as Fn
    from torcheasyrec_b200.engine import Pipeline

    with Fn.use_backend(OracleKernels()):
        monkeypatch.setenv("TZK_DIN_JAGGED", "0")
        pad = Pipeline("multi_tower_din_taobao", device="cpu", max_rows=200, seed=3)
        monkeypatch.setenv("TZK_DIN_JAGGED", "1")
        jag = Pipeline("multi_tower_din_taobao", device="cpu", max_rows=200, seed=3)
        jag.model.load_state_dict(pad.model.state_dict())
        assert not getattr(next(iter(pad.model.embedding_group.seq_emb_impls.values())), "_jagged_for_attention", None)
        assert next(iter(jag.model.embedding_group.seq_emb_impls.values()))._jagged_for_attention
        batch = pad.synthetic_batch(40, seed=9)
        with torch.no_grad():
            a, b = pad.model.predict(batch), jag.model.predict(batch)
        np.testing.assert_allclose(b["logits"].numpy(), a["logits"].numpy(), rtol=1e-5, atol=1e-6)
        la, lb = pad.eager_step(batch), jag.eager_step(batch)
        np.testing.assert_allclose(float(lb), float(la), rtol=1e-6)
        for ca, cb in zip(pad.model.sparse_collections(), jag.model.sparse_collections()):
            np.testing.assert_allclose(cb.weights.detach().numpy(), ca.weights.detach().numpy(), rtol=1e-5, atol=1e-7)
        for (n, pa), (_, pb) in zip(pad.model.named_parameters(), jag.model.named_parameters()):
            np.testing.assert_allclose(pb.detach().numpy(), pa.detach().numpy(), rtol=2e-4, atol=2e-6, err_msg=n)
