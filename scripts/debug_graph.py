"""Localises differences between a CUDA-graph step and an eager step (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torcheasyrec_b200.engine import GraphedTrainStep, Pipeline

a = Pipeline("dlrm_criteo", device="cuda:0", max_rows=5000, seed=3)
batches = [a.synthetic_batch(1024, seed=40 + i) for i in range(4)]
step = GraphedTrainStep(a, batches[0], warmup=3)
b = Pipeline("dlrm_criteo", device="cuda:0", max_rows=5000, seed=3)
b.model.load_state_dict(a.model.state_dict())
for ca, cb in zip(a.model.sparse_collections(), b.model.sparse_collections()):
    cb.opt_state.copy_(ca.opt_state)
import copy
b.dense_optimizer.load_state_dict(copy.deepcopy(a.dense_optimizer.state_dict()))


def report(tag):
    for (n, pa), (_, pb) in zip(a.model.named_parameters(), b.model.named_parameters()):
        d = (pa - pb).abs().max().item()
        print(f"{tag} param {n:60s} maxdiff {d:.3e} scale {pa.abs().max().item():.3e}")
    for ca, cb in zip(a.model.sparse_collections(), b.model.sparse_collections()):
        print(f"{tag} opt_state maxdiff {(ca.opt_state - cb.opt_state).abs().max().item():.3e}")
    sa, sb = a.dense_optimizer.state_dict()["state"], b.dense_optimizer.state_dict()["state"]
    for k in sa:
        for kk in sa[k]:
            va, vb = sa[k][kk], sb[k][kk]
            print(f"{tag} adam[{k}].{kk} maxdiff {(va.float() - vb.float()).abs().max().item():.3e} val {va.float().flatten()[0].item():.4e}")


report("before")
for i, bt in enumerate(batches[1:3]):
    step.load(bt.pin_memory())
    la = float(step.replay())
    lb = float(b.eager_step(bt.to("cuda:0")))
    print("step", i, "loss graph", la, "eager", lb)
    report(f"after{i}")
