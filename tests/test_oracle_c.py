"""The C/OpenMP restatement (CPU-baseline arm) against the pinned numpy oracle, end to end through the models."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_backend import OracleKernels  # noqa: E402

from oracle import build_c  # noqa: E402
from torcheasyrec_b200 import functional as Fn  # noqa: E402
from torcheasyrec_b200.engine import Pipeline  # noqa: E402


@pytest.mark.parametrize("name", ["dlrm_criteo", "deepfm_criteo"])
def test_c_oracle_matches_numpy_oracle(name):
    build_c.build()
    a = Pipeline(name, device="cpu", max_rows=500, seed=2)
    b = Pipeline(name, device="cpu", max_rows=500, seed=2)
    b.model.load_state_dict(a.model.state_dict())
    c_backend = OracleKernels(use_c=True)
    assert c_backend.use_c
    for it in range(2):
        batch = a.synthetic_batch(300, seed=it)
        with Fn.use_backend(OracleKernels()):
            la = float(a.eager_step(batch))
        with Fn.use_backend(c_backend):
            lb = float(b.eager_step(batch))
        np.testing.assert_allclose(la, lb, rtol=1e-6)
    for ca, cb in zip(a.model.sparse_collections(), b.model.sparse_collections()):
        np.testing.assert_allclose(ca.weights.detach().numpy(), cb.weights.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ca.opt_state.numpy(), cb.opt_state.numpy(), rtol=1e-4, atol=1e-12)
