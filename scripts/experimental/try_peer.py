"""First contact of the peer-memory sparse step (tzk_peer.cu + peer_exchange.py) with hardware — needs >= 2 GPUs.

    timeout 600 torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        scripts/experimental/try_peer.py [local_batch]

Two sharded copies of DLRM-Criteo seeded from the same unsharded weights: (a) the NCCL static-capacity exchange that
bench.py runs today, (b) the same collection switched to peer memory.  Checks, in this order (each prints PASS/FAIL):
  1. the barrier kernel alone (1000 rounds, must not hang — ALWAYS run under `timeout`),
  2. logits of (b) == logits of (a), bit for bit, before any update,
  3. after two training steps: every gathered table of (b) == (a), bit for bit; losses equal,
  4. CUDA-event time of 20 eager steps of each (max over ranks).
"""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)


def main() -> None:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = f"cuda:{local}"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))

    from torcheasyrec_b200.peer_exchange import enable_peer_exchange

    from torcheasyrec_b200.distributed import DenseGradSync, shard_model
    from torcheasyrec_b200.engine import Pipeline
    from torcheasyrec_b200.rank_models import dense_optimizer_from_config

    def say(msg):
        if rank == 0:
            print(msg, flush=True)

    def build(peer: bool):
        src = Pipeline("dlrm_criteo", device=dev, max_rows=200000, seed=5, capturable=False)
        p = Pipeline("dlrm_criteo", device=dev, max_rows=200000, seed=5, capturable=False)
        p.model.load_state_dict(src.model.state_dict())
        sharded = shard_model(p.model, dev, default="row_wise", source=src.model, static_capacity=2.5)
        p.model.set_sparse_optimizer(src.model.sparse_collections()[0].optimizer)
        p.dense_optimizer = dense_optimizer_from_config(p.cfg.train_config, p.model.dense_parameters())
        p.grad_sync = DenseGradSync(p.model.dense_parameters())
        states = []
        if peer:
            for sm in sharded:
                states += enable_peer_exchange(sm, B)
        return p, sharded, states

    a, sh_a, _ = build(False)
    b, sh_b, states = build(True)

    # 1. barrier
    for _ in range(1000):
        states[0].barrier()
    torch.cuda.synchronize()
    say(f"1. barrier x1000: PASS (epoch {int(states[0].epoch.item())})")

    batch = a.synthetic_batch(B, seed=77 + rank).to(dev)
    # 2. forward parity
    with torch.no_grad():
        pa, pb = a.model.predict(batch), b.model.predict(batch)
    ok = all(torch.equal(pa[k], pb[k]) for k in pa if k.startswith("logits"))
    say(f"2. logits bit-equal: {'PASS' if ok else 'FAIL'}")

    # 3. two steps, tables
    for _ in range(2):
        la, lb = a.eager_step(batch), b.eager_step(batch)
    for sm in sh_a + sh_b:
        sm.check_overflow()
    ok = float(la) == float(lb)
    for sa, sb in zip(sh_a, sh_b):
        for c in sa._configs:
            ok = ok and torch.equal(sa.gather_full_table(c.name), sb.gather_full_table(c.name))
    say(f"3. losses and every table bit-equal after 2 steps: {'PASS' if ok else 'FAIL'} (loss {float(la):.6f} / {float(lb):.6f})")

    # 4. time
    def time_steps(p, n=20):
        for _ in range(3):
            p.eager_step(batch)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            p.eager_step(batch)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ta, tb = time_steps(a), time_steps(b)
    say(f"4. eager step, local batch {B}, W={world}: NCCL static {ta:.3f} ms, peer memory {tb:.3f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
