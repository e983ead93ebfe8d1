"""The wide tower layer on the tcgen05 kernels INSIDE the model: two identically seeded DLRM pipelines, one stepping
with TZK_GEMM3X=0 (cuBLASLt BF16x9), one with TZK_GEMM3X=1, same batches; losses, logits and the wide layer's weights
after three steps must agree to fp32-GEMM tolerance.  Also captures the step in a CUDA graph with the switch on.

    timeout 300 python scripts/experimental/try_gemm3x_model.py [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    from torcheasyrec_b200 import dense_gemm as G
    from torcheasyrec_b200.engine import GraphedTrainStep, Pipeline

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    assert G._gemm3x_lib() is not None, "libtzk_gemm3x.so missing"
    pipes = {}
    for tag in ("0", "1"):
        pipes[tag] = Pipeline("dlrm_criteo", device="cuda", max_rows=100000, seed=5, capturable=False)
    batches = [pipes["0"].synthetic_batch(B, seed=11 + i).to("cuda") for i in range(3)]
    losses = {"0": [], "1": []}
    for tag, p in pipes.items():
        os.environ["TZK_GEMM3X"] = tag
        for b in batches:
            losses[tag].append(float(p.eager_step(b)))
        with torch.no_grad():
            p.logits = next(v for k, v in p.model.predict(batches[0]).items() if k.startswith("logits")).clone()
    torch.cuda.synchronize()
    print("losses cuBLASLt:", [f"{v:.7f}" for v in losses["0"]])
    print("losses tcgen05 :", [f"{v:.7f}" for v in losses["1"]])
    dl = max(abs(a - b) for a, b in zip(losses["0"], losses["1"]))
    dlog = (pipes["0"].logits - pipes["1"].logits).abs().max().item()
    dw = max((a - b).abs().max().item() for (n, a), (_, b) in zip(pipes["0"].model.named_parameters(),
                                                                  pipes["1"].model.named_parameters())
             if not n.endswith("weights"))
    print(f"max |loss diff| {dl:.2e}, max |logit diff| {dlog:.2e}, max dense-parameter diff {dw:.2e}: "
          f"{'PASS' if dl < 1e-5 and dlog < 1e-4 and dw < 1e-4 else 'FAIL'}", flush=True)
    # the switch inside a captured step
    os.environ["TZK_GEMM3X"] = "1"
    cap = Pipeline("dlrm_criteo", device="cuda", max_rows=100000, seed=5, capturable=True)
    step = GraphedTrainStep(cap, batches[0])
    for b in batches:
        step.load(b)
        last = float(step.replay())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        step.load(batches[i % 3])
        step.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"captured step with TZK_GEMM3X=1: loss {last:.6f}, {e0.elapsed_time(e1) / 20:.3f} ms/step at B={B}", flush=True)


if __name__ == "__main__":
    main()
