"""Pins the [EXT] half of the oracle (bucketize, KJT layout / permute, padded-dense, regroup, row-wise shard geometry)
and the package's own KJT / KeyedTensor / JaggedTensor classes to vectors restated from the upstream projects' unit
tests and docstrings (tests/golden/ext_vectors.py: symbol + derivation per vector).  GPU: the same vectors through the
CUDA kernels."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import ext_vectors as V  # noqa: E402
from oracle import tzk_oracle as O  # noqa: E402
from oracle_backend import OracleKernels  # noqa: E402

from torcheasyrec_b200 import functional as Fn  # noqa: E402
from torcheasyrec_b200.distributed import local_rows, make_plan, rw_block  # noqa: E402
from torcheasyrec_b200.embedding_modules import EmbeddingBagConfig  # noqa: E402
from torcheasyrec_b200.sparse import JaggedTensor, KeyedJaggedTensor, KeyedTensor  # noqa: E402


def _bucketize_case():
    v = V.FBGEMM_BLOCK_BUCKETIZE
    ids = np.asarray(v["indices"], dtype=np.int64)
    off = O.lengths_to_offsets(np.asarray(v["lengths"], dtype=np.int32))
    return v, ids, off


def test_oracle_bucketize_matches_fbgemm_block_bucketize_test_vector():
    v, ids, off = _bucketize_case()
    ol, oo, oi, op = O.bucketize_rw(ids, off, v["T"], v["B"], v["my_size"], v["block_sizes"])
    assert ol.tolist() == v["new_lengths"]
    assert oi.tolist() == v["new_indices"]
    assert oo.tolist() == np.concatenate([[0], np.cumsum(v["new_lengths"])]).tolist()
    inv = np.empty(len(op), dtype=np.int64)
    inv[op] = np.arange(len(op))            # op[new] = old position  ->  unbucketize_permute[old] = new
    assert inv.tolist() == v["unbucketize_permute"]


def test_peer_bucketize_model_agrees_with_the_fbgemm_vector_per_destination():
    """tzk_peer_bucketize's wire layout (destination-major, (feature, bag, position) order inside a destination) holds
    the same ids in the same order as fbgemm's [bucket][t][b] output."""
    v, ids, off = _bucketize_case()
    T, B, W = v["T"], v["B"], v["my_size"]
    cap = 16
    wk, wi, cnt = torch.zeros(W * cap, dtype=torch.int64), torch.zeros(W * cap, dtype=torch.int32), torch.zeros(W + 1, dtype=torch.int32)
    OracleKernels().peer_bucketize(torch.from_numpy(ids), torch.from_numpy(off), T, B, W,
                                   torch.tensor(v["block_sizes"]), torch.zeros(T, dtype=torch.int32),
                                   torch.tensor([1000] * T), torch.zeros(W * T, dtype=torch.int64), True, cap, wk, wi, cnt)
    per_dest = np.add.reduceat(np.asarray(v["new_lengths"]), np.arange(0, W * T * B, T * B))
    assert cnt[:W].tolist() == per_dest.tolist() and int(cnt[W]) == 0
    got = np.concatenate([wk[r * cap:r * cap + per_dest[r]].numpy() for r in range(W)])
    assert got.tolist() == v["new_indices"]


def test_kjt_layout_and_permute_match_the_torchrec_docstring():
    v = V.TORCHREC_KJT_DOC
    kjt = KeyedJaggedTensor.from_lengths_sync(v["keys"], torch.tensor(v["values"]), torch.tensor(v["lengths"], dtype=torch.int32))
    assert kjt.stride() == v["stride"]
    assert kjt.offsets().tolist() == v["offsets"]
    assert kjt.length_per_key() == v["length_per_key"]
    assert O.lengths_to_offsets(np.asarray(v["lengths"], np.int32)).tolist() == v["offsets"]
    ids, lens = O.kjt_permute(np.asarray(v["values"], np.int64), np.asarray(v["lengths"], np.int32), [1, 0], v["stride"])
    assert ids.tolist() == v["permuted_values"] and lens.tolist() == v["permuted_lengths"]
    with Fn.use_backend(OracleKernels()):
        p = kjt.permute([1, 0])
    assert p.keys() == v["permuted_keys"] and p.values().tolist() == v["permuted_values"]
    assert p.lengths().tolist() == v["permuted_lengths"]
    d = kjt.to_dict()
    assert d["Feature1"].values().tolist() == [3, 4, 5, 6, 7] and d["Feature0"].lengths().tolist() == [2, 0, 1]


def test_to_padded_dense_matches_the_torchrec_docstring():
    v = V.TORCHREC_TO_PADDED_DENSE_DOC
    vals, off = np.asarray(v["values"], np.float32), np.asarray(v["offsets"], np.int64)
    assert O.to_padded_dense(vals[:, None], off, 3)[:, :, 0].tolist() == v["dense_default"]
    with Fn.use_backend(OracleKernels()):
        jt = JaggedTensor(torch.tensor(v["values"]), offsets=torch.tensor(v["offsets"]))
        assert jt.to_padded_dense().tolist() == v["dense_default"]
        assert jt.to_padded_dense(desired_length=2, padding_value=10.0).tolist() == v["dense_len2_pad10"]


def test_keyed_tensor_and_regroup_match_the_torchrec_docstring():
    v = V.TORCHREC_KEYED_TENSOR_DOC
    kt = KeyedTensor(v["keys"], v["length_per_key"], torch.tensor(v["values"]))
    assert kt.offset_per_key() == v["offset_per_key"]
    assert kt.to_dict()["Embedding B"].tolist() == v["embedding_b"]
    got = O.regroup([(v["keys"], v["length_per_key"], np.asarray(v["values"], np.float32))], v["regroup_groups"])
    assert [g.tolist() for g in got] == v["regrouped"]
    with Fn.use_backend(OracleKernels()):
        outs = KeyedTensor.regroup([kt], v["regroup_groups"])
    assert [o.tolist() for o in outs] == v["regrouped"]


@pytest.mark.parametrize("rows,W,block,per_rank", V.TORCHREC_RW_GEOMETRY)
def test_row_wise_shard_geometry(rows, W, block, per_rank):
    assert O.rw_block_size(rows, W) == block == rw_block(rows, W)
    cfg = EmbeddingBagConfig(num_embeddings=rows, embedding_dim=16, name="t", feature_names=["f"])
    plan = make_plan([cfg], W, "row_wise")
    assert [local_rows(cfg, plan["t"], r) for r in range(W)] == per_rank
    assert sum(per_rank) == rows
    # every id lands on the rank that owns its row (bucketize rule = shard geometry), also for the last, short shard
    ids = np.unique(np.concatenate([np.arange(min(rows, 50)), np.arange(max(rows - 50, 0), rows),
                                    np.arange(0, rows, max(rows // 97, 1))])).astype(np.int64)
    off = np.arange(len(ids) + 1, dtype=np.int64)
    ol, oo, oi, op = O.bucketize_rw(ids, off, 1, len(ids), W, [block])
    dest_start = oo[::len(ids)]
    for r in range(W):
        mine = oi[dest_start[r]:dest_start[r + 1]]
        assert (mine >= 0).all() and (mine < max(per_rank[r], 1)).all() and (len(mine) == 0 or per_rank[r] > 0)
        assert (ids[op[dest_start[r]:dest_start[r + 1]]] == mine + r * block).all()


def test_embedding_bag_documentation_shape():
    v = V.EMBEDDING_BAG_DOC_SHAPE
    W = np.repeat(np.arange(v["rows"], dtype=np.float32)[:, None], v["dim"], axis=1)
    ids = np.asarray(v["input"], np.int64)
    off = np.asarray(v["offsets"] + [len(ids)], np.int64)
    for mode, pool in (("sum", 0), ("mean", 1)):
        got = O.pooled_lookup([W], [0], [pool], ids, off, 2)
        want = torch.nn.functional.embedding_bag(torch.from_numpy(ids), torch.from_numpy(W), torch.tensor(v["offsets"]), mode=mode)
        assert got.tolist() == v[mode] == want.tolist()


# ---- the same vectors through the CUDA kernels --------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_kernels_match_the_ext_vectors(kernels):
    v, ids, off = _bucketize_case()
    dev = "cuda"
    T, B, W = v["T"], v["B"], v["my_size"]
    ol, oo, oi, _, inv = kernels.bucketize_rw(torch.from_numpy(ids).to(dev), torch.from_numpy(off).to(dev), T, B, W,
                                              torch.tensor(v["block_sizes"], device=dev), want_inv=True)
    assert ol.tolist() == v["new_lengths"] and oi.tolist() == v["new_indices"]
    assert inv.tolist() == v["unbucketize_permute"]
    cap = 16
    wk = torch.zeros(W * cap, dtype=torch.int64, device=dev)
    wi = torch.zeros(W * cap, dtype=torch.int32, device=dev)
    cnt = torch.zeros(W + 1, dtype=torch.int32, device=dev)
    kernels.peer_bucketize(torch.from_numpy(ids).to(dev), torch.from_numpy(off).to(dev), T, B, W,
                           torch.tensor(v["block_sizes"], device=dev), torch.zeros(T, dtype=torch.int32, device=dev),
                           torch.tensor([1000] * T, device=dev), torch.zeros(W * T, dtype=torch.int64, device=dev), True,
                           cap, wk, wi, cnt)
    per_dest = np.add.reduceat(np.asarray(v["new_lengths"]), np.arange(0, W * T * B, T * B))
    assert cnt[:W].tolist() == per_dest.tolist()
    got = torch.cat([wk[r * cap:r * cap + int(per_dest[r])] for r in range(W)])
    assert got.tolist() == v["new_indices"]
    k = V.TORCHREC_KJT_DOC
    kjt = KeyedJaggedTensor.from_lengths_sync(k["keys"], torch.tensor(k["values"], device=dev),
                                              torch.tensor(k["lengths"], dtype=torch.int32, device=dev))
    assert kjt.offsets().tolist() == k["offsets"]
    p = kjt.permute([1, 0])
    assert p.values().tolist() == k["permuted_values"] and p.lengths().tolist() == k["permuted_lengths"]
    d = V.TORCHREC_TO_PADDED_DENSE_DOC
    jt = JaggedTensor(torch.tensor(d["values"], device=dev), offsets=torch.tensor(d["offsets"], device=dev))
    assert jt.to_padded_dense().tolist() == d["dense_default"]
    r = V.TORCHREC_KEYED_TENSOR_DOC
    kt = KeyedTensor(r["keys"], r["length_per_key"], torch.tensor(r["values"], device=dev))
    assert [o.tolist() for o in KeyedTensor.regroup([kt], r["regroup_groups"])] == r["regrouped"]
