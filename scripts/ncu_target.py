"""Target process for scripts/ncu_traffic.py: the pooled gather and the fused backward of the headline workload
(DLRM-Criteo, full hash sizes, B = 65536), launched stand-alone (no CUDA graph) so that ncu can replay each kernel.
Usage: python scripts/ncu_target.py [reps]   (the LAST repetition is the one worth reading: caches are warm-ish)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from torcheasyrec_b200.engine import Pipeline  # noqa: E402
from torcheasyrec_b200.kernels import default_kernels  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    B = 65536
    dev = torch.device("cuda", 0)
    pipe = Pipeline("dlrm_criteo", device=dev)
    kern = default_kernels()
    ebc = pipe.model.sparse_collections()[0]
    lay, spec = ebc.layout, ebc.optimizer
    grad = torch.randn((B, lay.total_dim), device=dev) * 1e-3
    out = torch.empty((B, lay.total_dim), device=dev)
    for i in range(reps):
        b = pipe.synthetic_batch(B, seed=20260923 + i).to(dev)
        kjt = ebc._select(b.sparse_features[sorted(b.sparse_features)[0]])
        off = kern.lengths_to_offsets(kjt.lengths())
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push(f"tzk_rep{i}")
        kern.pooled_gather_fwd(ebc.weights.data, lay, kjt.values(), off, B, out)
        kern.fused_bwd(spec.kind, True, grad, ebc.weights.data, ebc.opt_state, lay, kjt.values(), off, B, spec.lr,
                       spec.eps, 1.0)
        torch.cuda.nvtx.range_pop()
        torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
