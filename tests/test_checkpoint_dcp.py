"""checkpoints in the reference's layout (tzrec/utils/checkpoint_util.py:1109-1160 / :943-1107) —
`<dir>/model` + `<dir>/optimizer` written by torch.distributed.checkpoint with the reference's per-table key names,
`<dir>/plan` json — and their round trips across world sizes and plans (gloo on CPU, oracle backend as compute):
   W=2 (mixed table-wise + row-wise) save  ->  W=1 unsharded restore  ->  same tables / optimizer state / dense params
   W=1 save -> W=2 row-wise restore;  W=2 row-wise save -> W=2 table-wise restore (re-sharding by chunk metadata)."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from test_distributed_cpu import _free_port  # noqa: E402

NAME = "dlrm_criteo"


def _build(world, sharding, seed, rw_min_rows=0):
    from torcheasyrec_b200.distributed import DenseGradSync, shard_model
    from torcheasyrec_b200.engine import Pipeline
    from torcheasyrec_b200.rank_models import dense_optimizer_from_config

    p = Pipeline(NAME, device="cpu", max_rows=300, seed=seed, capturable=False)
    if sharding is not None:
        ref = Pipeline(NAME, device="cpu", max_rows=300, seed=seed, capturable=False)
        shard_model(p.model, "cpu", default=sharding, rw_min_rows=rw_min_rows, source=ref.model)
        p.model.set_sparse_optimizer(ref.model.sparse_collections()[0].optimizer)
        p.dense_optimizer = dense_optimizer_from_config(p.cfg.train_config, p.model.dense_parameters())
        p.grad_sync = DenseGradSync(p.model.dense_parameters())
    return p


def _full_tables(p):
    """{table: full [rows, D] weight}, {table: full optimizer state} whatever the sharding."""
    from torcheasyrec_b200.checkpoint import _collections, _local_state
    from torcheasyrec_b200.distributed import TABLE_WISE, _ShardedBase

    w, s = {}, {}
    for _, m in _collections(p.model):
        if isinstance(m, _ShardedBase):
            for g in m.groups:
                for t, c in enumerate(g.configs):
                    w[c.name] = m.gather_full_table(c.name).clone()
                    sh = m.plan[c.name]
                    block = c.num_embeddings if sh.kind == TABLE_WISE else sh.block
                    pad = torch.zeros((block, c.embedding_dim))
                    loc = _local_state(g.local, t, "momentum1")
                    if loc is not None and loc.shape[0]:
                        pad[:loc.shape[0]] = loc
                    parts = [torch.empty_like(pad) for _ in range(m.world)]
                    dist.all_gather(parts, pad)
                    s[c.name] = (parts[sh.owner] if sh.kind == TABLE_WISE else torch.cat(parts))[:c.num_embeddings].clone()
        else:
            for t, c in enumerate(m._configs):
                w[c.name] = m.table_weight(t).clone()
                s[c.name] = _local_state(m, t, "momentum1").clone()
    return w, s


def _worker(rank, world, port, tmp, phase, sharding, rw_min_rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        from oracle_backend import OracleKernels

        from torcheasyrec_b200 import functional as Fn
        from torcheasyrec_b200.checkpoint import list_checkpoint_keys, restore_model, save_model

        with Fn.use_backend(OracleKernels()):
            if phase == "save":
                p = _build(world, sharding, seed=5, rw_min_rows=rw_min_rows)
                for i in range(2):          # two steps: non-trivial Adagrad / Adam state
                    p.eager_step(p.synthetic_batch(32, seed=10 + i + 100 * rank))
                save_model(tmp, p.model, p.dense_optimizer)
                w, s = _full_tables(p)
                if rank == 0:
                    torch.save({"w": w, "s": s, "dense": {n: v.detach().clone() for n, v in p.model.named_parameters()
                                                          if not n.endswith(".weights")},
                                "adam": {k: v.clone() for k, v in _dense_state(p).items()}}, os.path.join(tmp, "expect.pt"))
                    keys = list_checkpoint_keys(tmp)
                    assert not any("shards" in k or k.endswith(".weights") for k in keys), keys
                    pre = "embedding_group.emb_impls.__BASE__.ebc.embedding_bags."
                    assert f"{pre}cat_0_emb.weight" in keys
                    assert f"state.{pre}cat_0_emb.weight.cat_0_emb.momentum1" in keys
                    assert any(k.startswith("state.") and k.endswith(".exp_avg_sq") for k in keys)
                    if sharding is not None:
                        plan = json.load(open(os.path.join(tmp, "plan")))
                        mod = plan["embedding_group.emb_impls.__BASE__.ebc"]
                        assert set(mod["cat_0_emb"]) == {"sharding_type", "compute_kernel", "ranks"}
            else:
                p = _build(world, sharding, seed=99, rw_min_rows=rw_min_rows)     # different init: everything must come from disk
                p.eager_step(p.synthetic_batch(32, seed=1 + rank))               # materialises the dense optimizer state
                restore_model(tmp, p.model, p.dense_optimizer)
                exp = torch.load(os.path.join(tmp, "expect.pt"))
                w, s = _full_tables(p)
                for k in exp["w"]:
                    assert torch.equal(w[k], exp["w"][k]), k
                    assert torch.equal(s[k], exp["s"][k]), k
                for n, v in p.model.named_parameters():
                    if not n.endswith(".weights"):
                        assert torch.equal(v.detach(), exp["dense"][n]), n
                for k, v in _dense_state(p).items():
                    assert torch.equal(v, exp["adam"][k]), k
                # and the restored model keeps training (fused state is live)
                p.eager_step(p.synthetic_batch(32, seed=3 + rank))
        q.put((rank, "ok"))
    except Exception:
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _dense_state(p):
    from torcheasyrec_b200.checkpoint import dense_optimizer_state_dict

    return dense_optimizer_state_dict(p.model, p.dense_optimizer)


def _run(world, tmp, phase, sharding, rw_min_rows=0):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, tmp, phase, sharding, rw_min_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    bad = [(r, m) for r, m in res if m != "ok"]
    assert not bad, "\n".join(f"rank {r}: {m}" for r, m in bad)


@pytest.mark.parametrize("save_cfg,load_cfg", [((2, "mixed", 200), (1, None, 0)), ((1, None, 0), (2, "row_wise", 0)),
                                               ((2, "row_wise", 0), (2, "table_wise", 0))])
def test_checkpoint_round_trip_across_world_sizes_and_plans(tmp_path, save_cfg, load_cfg):
    tmp = str(tmp_path)
    _run(save_cfg[0], tmp, "save", save_cfg[1], save_cfg[2])
    assert os.path.exists(os.path.join(tmp, "model", ".metadata")) and os.path.exists(os.path.join(tmp, "optimizer", ".metadata"))
    _run(load_cfg[0], tmp, "load", load_cfg[1], load_cfg[2])
