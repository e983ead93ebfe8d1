"""`DataParser.to_batch` (reference semantics: tzrec/datasets/data_parser.py:402-594) and the
final-layout build `to_batch_into` (one pinned arena, one H2D copy) produce the same Batch; the flattening example in
the comments of tzrec/datasets/utils.py:311-321 is replayed."""
import numpy as np
import torch

from torcheasyrec_b200.data_parser import DataParser
from torcheasyrec_b200.engine import Pipeline


def _columns(pipe, B, seed):
    """Per-feature columns as a reader would hand them over, taken apart from a synthetic batch."""
    b = pipe.synthetic_batch(B, seed=seed)
    cols = {}
    for dg, kjt in b.sparse_features.items():
        for k, jt in kjt.to_dict().items():
            cols[f"{k}.values"], cols[f"{k}.lengths"] = jt.values().clone(), jt.lengths().clone()
    for dg, kt in b.dense_features.items():
        for k, v in kt.to_dict().items():
            cols[f"{k}.values"] = v.clone()
    for k, v in b.labels.items():
        cols[k] = v.clone()
    return b, cols


def _same(a, b):
    assert set(a.sparse_features) == set(b.sparse_features) and set(a.dense_features) == set(b.dense_features)
    for dg in a.sparse_features:
        x, y = a.sparse_features[dg], b.sparse_features[dg]
        assert x.keys() == y.keys() and x.stride() == y.stride()
        assert torch.equal(x.values().cpu(), y.values().cpu()) and torch.equal(x.lengths().cpu(), y.lengths().cpu())
        assert x.length_per_key() == y.length_per_key()
    for dg in a.dense_features:
        assert a.dense_features[dg].keys() == b.dense_features[dg].keys()
        assert torch.equal(a.dense_features[dg].values().cpu(), b.dense_features[dg].values().cpu())
    for k in a.labels:
        assert torch.equal(a.labels[k].cpu(), b.labels[k].cpu())


def test_to_batch_and_final_layout_build_agree_with_the_generator():
    for name, B in (("dlrm_criteo", 64), ("multi_tower_din_taobao", 33), ("mmoe_taobao", 16)):
        pipe = Pipeline(name, device="cpu", max_rows=100)
        want, cols = _columns(pipe, B, seed=3)
        parser = DataParser(pipe.features, pipe.labels)
        _same(parser.to_batch(cols), want)
        arena = parser.make_arena(B, {dg: int(k.values().numel()) + 7 for dg, k in want.sparse_features.items()})
        host = parser.to_batch_into(cols, arena)
        _same(host, want)
        # second batch through the same arena: nothing of the first one leaks
        want2, cols2 = _columns(pipe, B, seed=4)
        _same(parser.to_batch_into(cols2, arena), want2)
        # every tensor of the host batch lives inside the one arena buffer
        lo, hi = arena.host.data_ptr(), arena.host.data_ptr() + arena.host.numel()
        for kjt in host.sparse_features.values():
            assert lo <= kjt.values().data_ptr() < hi and lo <= kjt.lengths().data_ptr() < hi


def test_kjt_flattening_example_of_the_reference():
    """tzrec/datasets/utils.py:311-321: features f1 = [[1,2],[3]], f2 = [[4],[5,6]] (B = 2) flatten key-major to
    values [1,2,3,4,5,6], lengths [2,1,1,2]."""
    class F:
        def __init__(self, name):
            self.name, self.data_group = name, "__BASE__"
            self.is_sequence, self.is_sparse, self.is_weighted, self.value_dim, self.stub_type = False, True, False, 0, False

    parser = DataParser([F("f1"), F("f2")], labels=["y"])
    cols = {"f1.values": torch.tensor([1, 2, 3]), "f1.lengths": torch.tensor([2, 1], dtype=torch.int32),
            "f2.values": torch.tensor([4, 5, 6]), "f2.lengths": torch.tensor([1, 2], dtype=torch.int32),
            "y": torch.tensor([1.0, 0.0])}
    b = parser.to_batch(cols)
    kjt = b.sparse_features["__BASE__"]
    assert kjt.keys() == ["f1", "f2"] and kjt.values().tolist() == [1, 2, 3, 4, 5, 6]
    assert kjt.lengths().tolist() == [2, 1, 1, 2] and kjt.stride() == 2 and kjt.length_per_key() == [3, 3]
    arena = parser.make_arena(2, {"__BASE__": 10})
    h = parser.to_batch_into(cols, arena)
    assert h.sparse_features["__BASE__"].values().tolist() == [1, 2, 3, 4, 5, 6]
    assert h.sparse_features["__BASE__"].lengths().tolist() == [2, 1, 1, 2]


def test_weighted_and_multi_value_sequence_columns():
    class F:
        def __init__(self, name, seq=False, weighted=False, value_dim=0):
            self.name, self.data_group = name, "__BASE__"
            self.is_sequence, self.is_sparse, self.is_weighted, self.value_dim, self.stub_type = seq, True, weighted, value_dim, False

    parser = DataParser([F("w", weighted=True), F("plain"), F("mv", seq=True, value_dim=0)])
    cols = {"w.values": torch.tensor([7, 8, 9]), "w.lengths": torch.tensor([1, 2], dtype=torch.int32),
            "w.weights": torch.tensor([0.5, 2.0, 3.0]),
            "plain.values": torch.tensor([1, 2]), "plain.lengths": torch.tensor([1, 1], dtype=torch.int32),
            # sample 0: 2 steps with 1 and 2 ids; sample 1: 1 step with 3 ids  (data_parser.py:556-566)
            "mv.values": torch.tensor([10, 11, 12, 13, 14, 15]), "mv.lengths": torch.tensor([2, 1], dtype=torch.int32),
            "mv.key_lengths": torch.tensor([1, 2, 3], dtype=torch.int32)}
    b = parser.to_batch(cols)
    kjt = b.sparse_features["__BASE__"]
    assert kjt.lengths().tolist() == [1, 2, 1, 1, 3, 3]
    assert kjt.weights_or_none().tolist() == [0.5, 2.0, 3.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]
    mv = b.sequence_mulval_lengths["__BASE__"]
    assert mv.keys() == ["mv"] and mv.values().tolist() == [1, 2, 3] and mv.lengths().tolist() == [2, 1]
    arena = parser.make_arena(2, {"__BASE__": 16})
    h = parser.to_batch_into(cols, arena)
    assert h.sparse_features["__BASE__"].lengths().tolist() == [1, 2, 1, 1, 3, 3]
    assert h.sparse_features["__BASE__"].weights_or_none().tolist() == kjt.weights_or_none().tolist()
