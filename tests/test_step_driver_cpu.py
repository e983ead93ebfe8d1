"""Host logic of the step driver that only shows on a GPU when it is wrong: the deferred join of the fused sparse update
(engine.Pipeline.step_body) and the gating of the fused tower tail (rank_models.DLRM._fused_tail) — exercised on CPU with
the oracle backend, where both must be exact no-ops."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_backend import OracleKernels  # noqa: E402

from torcheasyrec_b200 import functional as Fn  # noqa: E402
from torcheasyrec_b200.engine import Pipeline  # noqa: E402


def test_defer_join_flags_are_scoped_to_the_step_and_survive_exceptions(monkeypatch):
    p = Pipeline("dlrm_criteo", device="cpu", max_rows=50, seed=1)
    joiners = p._sparse_joiners()
    assert joiners and all(not j.defer_join for j in joiners)
    seen = []
    orig = p.train_wrapper.forward

    def spy(batch):
        seen.append([j.defer_join for j in joiners])
        return orig(batch)

    monkeypatch.setattr(p.train_wrapper, "forward", spy)
    with Fn.use_backend(OracleKernels()):
        p.eager_step(p.synthetic_batch(16, seed=0))
    assert seen == [[True] * len(joiners)] and all(not j.defer_join for j in joiners)   # set inside the step only
    assert all(j._pending_join is None for j in joiners)

    def boom(batch):
        raise RuntimeError("forward failed")

    monkeypatch.setattr(p.train_wrapper, "forward", boom)
    with pytest.raises(RuntimeError, match="forward failed"):
        p.eager_step(p.synthetic_batch(16, seed=1))
    assert all(not j.defer_join for j in joiners)                                        # reset on the way out
    monkeypatch.setenv("TZK_DEFER_JOIN", "0")
    monkeypatch.setattr(p.train_wrapper, "forward", spy)
    seen.clear()
    with Fn.use_backend(OracleKernels()):
        p.eager_step(p.synthetic_batch(16, seed=2))
    assert seen == [[False] * len(joiners)]


def test_fused_tail_stays_out_of_the_way_off_cuda_and_in_eval_mode():
    p = Pipeline("dlrm_criteo", device="cpu", max_rows=50, seed=1)
    b = p.synthetic_batch(16, seed=0)
    with Fn.use_backend(OracleKernels()):
        pred = p.model.predict(b)
        assert getattr(p.model, "_tail_loss", None) is None and set(pred) == {"logits", "probs"}
        loss = p.model.loss(pred, b)["binary_cross_entropy"]
        ref = torch.nn.functional.binary_cross_entropy_with_logits(pred["logits"], b.labels[p.labels[0]].float())
        assert torch.allclose(loss, ref)
        # a stale hand-over from an earlier predict() must not leak into a later loss()
        p.model._tail_loss = torch.tensor(123.0)
        pred = p.model.predict(b)
        assert p.model._tail_loss is None
