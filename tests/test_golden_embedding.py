"""Committed golden vectors of the embedding path (tests/golden/embedding_path.npz): the oracle (CPU, always) and
the CUDA kernels (-m gpu) must both reproduce torch.embedding_bag / torch.optim results on the reference's own KJT
examples and on a multi-hot batch with duplicates and empty bags."""
import os

import numpy as np
import pytest
import torch

from oracle import tzk_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "embedding_path.npz"))
CASES = ["kjt2", "kjt6", "multihot"]


def _load(tag):
    rows, dims = G[f"{tag}_rows"].tolist(), G[f"{tag}_dims"].tolist()
    values, lengths, grad = G[f"{tag}_values"], G[f"{tag}_lengths"], G[f"{tag}_grad"]
    B = len(lengths) // len(rows)
    tables = [G[f"{tag}_table{t}"].copy() for t in range(len(rows))]
    return rows, dims, values, lengths, O.lengths_to_offsets(lengths), grad, B, tables


@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("mode", ["sum", "mean"])
@pytest.mark.parametrize("opt", ["sgd", "adagrad"])
def test_oracle_reproduces_golden(tag, mode, opt):
    rows, dims, values, lengths, offsets, grad, B, tables = _load(tag)
    F = len(rows)
    pool = [O.POOL_SUM if mode == "sum" else O.POOL_MEAN] * F
    got = O.pooled_lookup(tables, list(range(F)), pool, values, offsets, B)
    np.testing.assert_allclose(got, G[f"{tag}_{mode}_pooled"], rtol=1e-6, atol=1e-7)
    states = [np.zeros_like(t) for t in tables]
    O.fused_update(O.OPT_SGD if opt == "sgd" else O.OPT_ADAGRAD, tables, states, list(range(F)), pool, values,
                   offsets, B, grad, 0.05, 1e-8)
    for t in range(F):
        np.testing.assert_allclose(tables[t], G[f"{tag}_{mode}_{opt}_table{t}"], rtol=2e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("mode", ["sum", "mean"])
@pytest.mark.parametrize("opt", ["sgd", "adagrad"])
def test_cuda_reproduces_golden(kernels, tag, mode, opt):
    from torcheasyrec_b200.kernels import build_layout

    rows, dims, values, lengths, offsets, grad, B, tables = _load(tag)
    F = len(rows)
    pool = [O.POOL_SUM if mode == "sum" else O.POOL_MEAN] * F
    lay = build_layout(rows, dims, list(range(F)), pool).to("cuda")
    arena = np.zeros(lay.arena_elems, np.float32)
    for f in range(F):
        arena[lay.w_off[f]:lay.w_off[f] + tables[f].size] = tables[f].ravel()
    arena = torch.from_numpy(arena).cuda()
    ids, offs = torch.from_numpy(values).cuda(), torch.from_numpy(offsets).cuda()
    got = kernels.pooled_gather_fwd(arena, lay, ids, offs, B).cpu().numpy()
    np.testing.assert_allclose(got, G[f"{tag}_{mode}_pooled"], rtol=1e-5, atol=1e-7)
    state = torch.zeros_like(arena)
    kernels.fused_bwd(O.OPT_SGD if opt == "sgd" else O.OPT_ADAGRAD, True, torch.from_numpy(grad).cuda(), arena, state,
                      lay, ids, offs, B, 0.05, 1e-8, 1.0)
    out = arena.cpu().numpy()
    for f in range(F):
        np.testing.assert_allclose(out[lay.w_off[f]:lay.w_off[f] + tables[f].size].reshape(tables[f].shape),
                                   G[f"{tag}_{mode}_{opt}_table{f}"], rtol=2e-5, atol=1e-6)
