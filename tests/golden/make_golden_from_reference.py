"""Generates golden vectors from the REFERENCE's own modules (run in the build container only).

The reference package cannot be imported as a whole (torchrec / fbgemm_gpu / pyfg are absent), but
tzrec/modules/fm.py, interaction.py, mlp.py and mmoe.py are plain PyTorch: they are loaded file by file
through stub parent packages, executed on seeded inputs, and their outputs (and autograd gradients) are
stored as small .npz fixtures.  /root/reference does not travel to the GPU box; the fixtures do.

    python tests/golden/make_golden_from_reference.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub_packages():
    for name in ["tzrec", "tzrec.modules", "tzrec.utils", "tzrec.models"]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, name.replace(".", "/"))]
        sys.modules[name] = m


def main():
    _stub_packages()
    from tzrec.modules.fm import FactorizationMachine  # tzrec/modules/fm.py:16
    from tzrec.modules.interaction import InteractionArch  # tzrec/modules/interaction.py:57
    from tzrec.modules.mlp import MLP  # tzrec/modules/mlp.py:86

    torch.manual_seed(20260923)
    out = {}
    # FM: [B, N, D] (deepfm_criteo: N=26, D=16) + odd shapes
    for tag, (B, N, D) in {"fm_criteo": (64, 26, 16), "fm_small": (5, 3, 4), "fm_wide": (7, 9, 32)}.items():
        x = torch.randn(B, N, D, requires_grad=True)
        y = FactorizationMachine()(x)
        dy = torch.randn_like(y)
        y.backward(dy)
        out[f"{tag}_x"], out[f"{tag}_y"], out[f"{tag}_dy"], out[f"{tag}_dx"] = (
            x.detach().numpy(), y.detach().numpy(), dy.numpy(), x.grad.numpy())
    # InteractionArch: DLRM-Criteo N=27, D=16 and edge shapes (N=2, N not multiple of 4, D=4/64)
    for tag, (B, N, D) in {"ia_criteo": (48, 27, 16), "ia_min": (3, 2, 4), "ia_odd": (9, 13, 8),
                           "ia_wide": (4, 33, 64)}.items():
        x = torch.randn(B, N, D, requires_grad=True)
        z = InteractionArch(N)(x)
        dz = torch.randn_like(z)
        z.backward(dz)
        out[f"{tag}_x"], out[f"{tag}_z"], out[f"{tag}_dz"], out[f"{tag}_dx"] = (
            x.detach().numpy(), z.detach().numpy(), dz.numpy(), x.grad.numpy())
    # DLRM predict glue (tzrec/models/dlrm.py:113-131) re-enacted with the reference modules
    B, Ns, D = 32, 26, 16
    dense_mlp = MLP(13, [64, 16])
    final_mlp = MLP(351 + 16 + Ns * D, [64, 32])
    head = torch.nn.Linear(32, 1)
    dense_in = torch.rand(B, 13)
    sparse = torch.randn(B, Ns * D, requires_grad=True)
    dense_feat = dense_mlp(dense_in)
    feat = torch.cat([dense_feat.unsqueeze(1), sparse.reshape(-1, Ns, D)], dim=1)
    inter = InteractionArch(Ns + 1)(feat)
    all_feat = torch.cat([inter, dense_feat, sparse], dim=-1)
    logits = head(final_mlp(all_feat)).squeeze(1)
    labels = (torch.rand(B) < 0.25).float()
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)
    loss.backward()
    out["dlrm_dense_in"], out["dlrm_sparse"], out["dlrm_labels"] = dense_in.numpy(), sparse.detach().numpy(), labels.numpy()
    out["dlrm_all_feat"], out["dlrm_logits"], out["dlrm_loss"] = all_feat.detach().numpy(), logits.detach().numpy(), loss.detach().numpy()
    out["dlrm_dsparse"] = sparse.grad.numpy()
    sd = {}
    for prefix, mod in [("dense_mlp", dense_mlp), ("final_mlp", final_mlp), ("output_mlp", head)]:
        for k, v in mod.state_dict().items():
            sd[f"dlrm_sd__{prefix}.{k}"] = v.numpy()
        for k, p in mod.named_parameters():
            sd[f"dlrm_grad__{prefix}.{k}"] = p.grad.numpy()
    out.update(sd)
    np.savez_compressed(os.path.join(HERE, "ref_dense_modules.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_dense_modules.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
