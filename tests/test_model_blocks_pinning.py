"""The model shells that call the hot path (DIN attention, MMoE, TaskTower) against golden vectors produced by the
REFERENCE's own modules (tzrec/modules/sequence.py:65-128, mmoe.py:21-77, task_tower.py:21-52), loaded file by file in
the build container by tests/golden/make_golden_from_reference.py (`blocks`).  Same parameter names, so the reference
state_dict loads unchanged; outputs and every gradient must agree to fp32 round-off."""
import os

import numpy as np
import pytest
import torch

from torcheasyrec_b200.rank_models import DINEncoder, MMoEModule, TaskTower

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_model_blocks.npz"))


def _load(mod, tag):
    sd = {k[len(tag) + 5:]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(tag + "_sd__")}
    assert set(sd) == set(mod.state_dict()), (sorted(sd), sorted(mod.state_dict()))
    mod.load_state_dict(sd)
    return mod


def _check_grads(mod, tag, rtol=2e-5, atol=2e-6):
    for k, p in mod.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), GOLD[f"{tag}_grad__{k}"], rtol=rtol, atol=atol, err_msg=k)


def t(name, grad=False):
    x = torch.from_numpy(GOLD[name].copy())
    return x.requires_grad_(True) if grad else x


def test_din_encoder_matches_reference_module():
    enc = _load(DINEncoder(sequence_dim=24, query_dim=24, input="seq", attn_mlp={"hidden_units": [48, 16]}), "din")
    q, seq = t("din_q", True), t("din_seq", True)
    y = enc({"seq.query": q, "seq.sequence": seq, "seq.sequence_length": t("din_len")})
    np.testing.assert_allclose(y.detach().numpy(), GOLD["din_y"], rtol=1e-5, atol=1e-6)
    y.backward(t("din_dy"))
    np.testing.assert_allclose(q.grad.numpy(), GOLD["din_dq"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(seq.grad.numpy(), GOLD["din_dseq"], rtol=2e-5, atol=2e-6)
    _check_grads(enc, "din")


def test_din_encoder_query_padding_and_max_seq_length():
    enc = _load(DINEncoder(sequence_dim=16, query_dim=8, input="s", attn_mlp={"hidden_units": [32]}, max_seq_length=4),
                "din2")
    q, seq = t("din2_q", True), t("din2_seq", True)
    y = enc({"s.query": q, "s.sequence": seq, "s.sequence_length": t("din2_len")})
    np.testing.assert_allclose(y.detach().numpy(), GOLD["din2_y"], rtol=1e-5, atol=1e-6)
    y.sum().backward()
    np.testing.assert_allclose(q.grad.numpy(), GOLD["din2_dq"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(seq.grad.numpy(), GOLD["din2_dseq"], rtol=2e-5, atol=2e-6)
    _check_grads(enc, "din2")


def test_mmoe_matches_reference_module():
    mm = _load(MMoEModule(in_features=40, expert_mlp={"hidden_units": [64, 32, 16]}, num_expert=3, num_task=2), "mmoe")
    x = t("mmoe_x", True)
    ys = mm(x)
    for i, y in enumerate(ys):
        np.testing.assert_allclose(y.detach().numpy(), GOLD[f"mmoe_y{i}"], rtol=1e-5, atol=1e-6)
    torch.autograd.backward(ys, [t(f"mmoe_dy{i}") for i in range(len(ys))])
    np.testing.assert_allclose(x.grad.numpy(), GOLD["mmoe_dx"], rtol=2e-5, atol=2e-6)
    _check_grads(mm, "mmoe")


def test_task_tower_matches_reference_module():
    tt = _load(TaskTower(16, 1, mlp={"hidden_units": [32, 16, 8]}), "tower")
    x = t("tower_x", True)
    y = tt(x)
    np.testing.assert_allclose(y.detach().numpy(), GOLD["tower_y"], rtol=1e-5, atol=1e-6)
    y.sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), GOLD["tower_dx"], rtol=2e-5, atol=2e-6)
    _check_grads(tt, "tower")


def _to_jagged(seq: np.ndarray, lens: np.ndarray):
    rows = np.concatenate([seq[b, :l] for b, l in enumerate(lens)] + [np.zeros((0, seq.shape[2]), np.float32)])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return rows, off


@pytest.mark.parametrize("tag,kw", [("din", dict(sequence_dim=24, query_dim=24, input="seq", attn_mlp={"hidden_units": [48, 16]})),
                                    ("din2", dict(sequence_dim=16, query_dim=8, input="s", attn_mlp={"hidden_units": [32]},
                                                  max_seq_length=4))])
def test_jagged_din_attention_matches_the_reference_module(tag, kw):
    """the same reference-generated vectors through the JAGGED path (rows [N, Ds] + offsets, no padded
    tensor): output, query / row gradients on the valid positions and every parameter gradient; padded positions have
    zero gradient in the reference (masked before the softmax) and do not exist here."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    from oracle_backend import OracleKernels

    from torcheasyrec_b200 import functional as Fn

    enc = _load(DINEncoder(**kw), tag)
    g = kw["input"]
    lens = GOLD[f"{tag}_len"].astype(np.int64)
    rows, off = _to_jagged(GOLD[f"{tag}_seq"], np.minimum(lens, GOLD[f"{tag}_seq"].shape[1]))
    q = t(f"{tag}_q", True)
    r = torch.from_numpy(rows.copy()).requires_grad_(True)
    with Fn.use_backend(OracleKernels()):
        y = enc({f"{g}.query": q, f"{g}.sequence": r, f"{g}.sequence_length": torch.from_numpy(lens),
                 f"{g}.sequence_offsets": torch.from_numpy(off)})
        # a sample WITHOUT rows: the reference's softmax over an all-masked row is uniform over the T padded rows, which
        # are zeros in the real pipeline (to_padded_dense) — i.e. a zero output, as here; the fixture filled the padding
        # with random numbers, so those samples are compared with zero instead of the fixture
        has = lens > 0
        np.testing.assert_allclose(y.detach().numpy()[has], GOLD[f"{tag}_y"][has], rtol=1e-5, atol=1e-6)
        assert not y.detach().numpy()[~has].any()
        if tag == "din":
            y.backward(t("din_dy"))
        else:
            y.sum().backward()
    np.testing.assert_allclose(q.grad.numpy(), GOLD[f"{tag}_dq"], rtol=2e-5, atol=2e-6)
    want, _ = _to_jagged(GOLD[f"{tag}_dseq"], np.minimum(lens, GOLD[f"{tag}_seq"].shape[1]))
    np.testing.assert_allclose(r.grad.numpy(), want, rtol=2e-5, atol=2e-6)
    _check_grads(enc, tag)
