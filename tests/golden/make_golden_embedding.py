"""Golden vectors for the embedding path (run in the build container; fixtures travel to the GPU box).

The reference pins only shapes here, and torchrec/fbgemm cannot run, so the vectors are produced
by the pieces of the reference stack that DO run: torch.nn.functional.embedding_bag (what the un-sharded torchrec
EmbeddingBagCollection calls) for the pooled lookups and torch.optim.{SGD,Adagrad} on its dense autograd gradient for
the fused update.  Inputs are the KJT examples of the reference's own tests:
  * tzrec/modules/embedding_test.py:241-248  values=[1..7], lengths=[1,2,1,3], keys cat_a,cat_b, B=2
  * tzrec/modules/embedding_test.py:378-392  values=range(24), lengths=[1,1,1,1,3,3,3,3,2,2,2,2], 6 keys, B=2
plus a seeded multi-hot batch with L in {0,1,2,7,33}, duplicates inside a bag and across samples (§8c (3),(4)).

    python tests/golden/make_golden_embedding.py
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def run_case(tag, rows, dims, values, lengths, B, rng, out):
    F = len(rows)
    offsets = np.zeros(F * B + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    tables = [(rng.standard_normal((r, d)) * 0.1).astype(np.float32) for r, d in zip(rows, dims)]
    grad = rng.standard_normal((B, sum(dims))).astype(np.float32)
    out[f"{tag}_values"], out[f"{tag}_lengths"] = values.astype(np.int64), lengths.astype(np.int32)
    out[f"{tag}_rows"], out[f"{tag}_dims"], out[f"{tag}_grad"] = np.asarray(rows), np.asarray(dims), grad
    for t, w in enumerate(tables):
        out[f"{tag}_table{t}"] = w
    for mode in ("sum", "mean"):
        for opt in ("sgd", "adagrad"):
            params = [torch.nn.Parameter(torch.from_numpy(w.copy())) for w in tables]
            pooled = []
            for f in range(F):
                s, e = offsets[f * B], offsets[(f + 1) * B]
                pooled.append(torch.nn.functional.embedding_bag(
                    torch.from_numpy(values[s:e].astype(np.int64)), params[f],
                    torch.from_numpy(offsets[f * B:(f + 1) * B] - s), mode=mode))
            kt = torch.cat(pooled, dim=1)
            kt.backward(torch.from_numpy(grad))
            optim = (torch.optim.SGD(params, lr=0.05) if opt == "sgd" else
                     torch.optim.Adagrad(params, lr=0.05, eps=1e-8, initial_accumulator_value=0.0))
            optim.step()
            out[f"{tag}_{mode}_pooled"] = kt.detach().numpy()
            for t, p in enumerate(params):
                out[f"{tag}_{mode}_{opt}_table{t}"] = p.detach().numpy()


def main():
    rng = np.random.default_rng(20260923)
    out = {}
    run_case("kjt2", [10, 10], [16, 8], np.arange(1, 8), np.array([1, 2, 1, 3]), 2, rng, out)
    run_case("kjt6", [30] * 6, [16] * 6, np.arange(24), np.array([1, 1, 1, 1, 3, 3, 3, 3, 2, 2, 2, 2]), 2, rng, out)
    B, F = 12, 3
    lens = rng.choice([0, 1, 2, 7, 33], size=F * B).astype(np.int64)
    vals = np.concatenate([rng.integers(0, 9, size=int(n)) for n in lens]).astype(np.int64)   # 9 rows: many duplicates
    run_case("multihot", [9, 9, 9], [16, 4, 32], vals, lens, B, rng, out)
    np.savez_compressed(os.path.join(HERE, "embedding_path.npz"), **out)
    print("wrote embedding_path.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
