"""N>1 host logic on CPU: world_size-2 (and 3) gloo processes drive the sharded collections (bucketize ->
all-to-all -> owner gather -> return all-to-all -> pool; fused backward on the owners) with the oracle backend
as compute, and compare with the UNSHARDED model stepped on the concatenated global batch:
pooled outputs / logits bit-equal (W-invariance), updated tables and dense weights equal within 1e-5."""
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


def _free_port():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _concat_batches(batches):
    from torcheasyrec_b200.batch import Batch
    from torcheasyrec_b200.sparse import KeyedJaggedTensor, KeyedTensor

    out = Batch()
    b0 = batches[0]
    for dg, kjt0 in b0.sparse_features.items():
        F, vals, lens = len(kjt0.keys()), [], []
        dicts = [b.sparse_features[dg].to_dict() for b in batches]
        for k in kjt0.keys():
            for d in dicts:
                vals.append(d[k].values())
                lens.append(d[k].lengths())
        out.sparse_features[dg] = KeyedJaggedTensor(kjt0.keys(), torch.cat(vals), lengths=torch.cat(lens),
                                                    stride=sum(b.sparse_features[dg].stride() for b in batches))
    for dg, kt0 in b0.dense_features.items():
        out.dense_features[dg] = KeyedTensor(kt0.keys(), kt0.length_per_key(),
                                             torch.cat([b.dense_features[dg].values() for b in batches]))
    for k in b0.labels:
        out.labels[k] = torch.cat([b.labels[k] for b in batches])
    return out


def _worker(rank, world, port, name, sharding, rw_min_rows, result_q, use_cuda=False, static_capacity=None,
            sparse_opt=None, exchange="nccl"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = f"cuda:{rank}" if use_cuda else "cpu"
    if use_cuda:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
    try:
        import contextlib

        from oracle_backend import OracleKernels

        from torcheasyrec_b200 import functional as Fn
        from torcheasyrec_b200.distributed import DenseGradSync, shard_model
        from torcheasyrec_b200.engine import Pipeline

        B = 48
        # CPU: host logic over gloo with the oracle as compute; GPU: the CUDA kernels over NCCL
        with (contextlib.nullcontext() if use_cuda else Fn.use_backend(OracleKernels())):
            ref = Pipeline(name, device=dev, max_rows=300, seed=5, capturable=False)   # unsharded twin
            shd = Pipeline(name, device=dev, max_rows=300, seed=5, capturable=False)
            shd.model.load_state_dict(ref.model.state_dict())
            sharded = shard_model(shd.model, dev, default=sharding, rw_min_rows=rw_min_rows, source=ref.model,
                                  static_capacity=static_capacity, exchange=exchange)
            if sparse_opt is not None:     # e.g. "adam": second state + device-side step counter on every shard
                from torcheasyrec_b200.embedding_modules import SparseOptimizerSpec

                ref.model.set_sparse_optimizer(SparseOptimizerSpec.from_name(sparse_opt, lr=0.01))
            shd.model.set_sparse_optimizer(ref.model.sparse_collections()[0].optimizer)
            from torcheasyrec_b200.rank_models import dense_optimizer_from_config

            shd.dense_optimizer = dense_optimizer_from_config(shd.cfg.train_config, shd.model.dense_parameters())
            shd.grad_sync = DenseGradSync(shd.model.dense_parameters())
            batches = [ref.synthetic_batch(B, seed=77 + r) for r in range(world)]
            glob = _concat_batches(batches).to(dev)
            batches = [b.to(dev) for b in batches]
            # forward parity (before any update)
            with torch.no_grad():
                p_ref = ref.model.predict(glob)
                p_shd = shd.model.predict(batches[rank])
            for k, v in p_shd.items():
                if k.startswith("logits"):
                    want = p_ref[k][rank * B:(rank + 1) * B].cpu().numpy()
                    if use_cuda:   # cuBLAS may pick another kernel for batch B vs W*B: not bit-stable across shapes
                        np.testing.assert_allclose(v.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
                    else:
                        np.testing.assert_array_equal(v.cpu().numpy(), want)
            # one train step on both
            for _ in range(2):
                loss_ref = ref.eager_step(glob)
                loss_shd = shd.eager_step(batches[rank])
            for sm in sharded:
                sm.check_overflow()
            t = torch.tensor([float(loss_shd)], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            np.testing.assert_allclose(t.item() / world, float(loss_ref), rtol=1e-6)
            # updated tables: gather shards, compare with the unsharded twin
            ref_tables = {}
            for coll in ref.model.sparse_collections():
                for ti, c in enumerate(coll._configs):
                    ref_tables[(type(coll).__name__, c.name)] = coll.table_weight(ti)
            for sm in sharded:
                kind = "EmbeddingBagCollection" if sm._pooled else "EmbeddingCollection"
                for c in sm._configs:
                    full = sm.gather_full_table(c.name)
                    # dL/dlogit is 1/B per rank then /W on the owners vs 1/(W*B) in the twin: same value, one more
                    # fp32 rounding per contribution -> a few ulp after two Adagrad steps
                    np.testing.assert_allclose(full.cpu().numpy(), ref_tables[(kind, c.name)].cpu().numpy(), rtol=5e-5, atol=1e-6,
                                               err_msg=f"{kind}.{c.name}")
            dense = lambda m: sorted((n, p) for n, p in m.named_parameters() if not n.endswith("weights"))
            for (n1, p1), (n2, p2) in zip(dense(ref.model), dense(shd.model)):
                assert n1 == n2
                # Adam normalises by sqrt(v): tiny gradient differences (mean over 2B vs mean of two means) are amplified
                np.testing.assert_allclose(p2.detach().cpu().numpy(), p1.detach().cpu().numpy(), rtol=2e-4, atol=2e-6, err_msg=n1)
        result_q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        import traceback

        result_q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _run(world, name, sharding, rw_min_rows=0, use_cuda=False, static_capacity=None, sparse_opt=None, exchange="nccl"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, sharding, rw_min_rows, q, use_cuda, static_capacity,
                                               sparse_opt, exchange))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    bad = [r for r in results if r[1] != "ok"]
    assert not bad, "\n".join(f"rank {r}: {msg}" for r, msg in bad)


@pytest.mark.parametrize("sharding", ["row_wise", "table_wise"])
def test_dlrm_two_ranks(sharding):
    _run(2, "dlrm_criteo", sharding)


def test_dlrm_static_capacity_exchange_two_ranks():
    # fixed-shape (graph-capturable) exchange: padded wire slots, zero-gradient padding, same results
    _run(2, "dlrm_criteo", "row_wise", static_capacity=2.5)


def test_deepfm_mixed_three_ranks():
    # two dim groups (wide D=4, deep D=16); tables with >= 200 rows row-wise, the tiny ones table-wise
    _run(3, "deepfm_criteo", "mixed", rw_min_rows=200)


@pytest.mark.parametrize("world,name,sharding,rw_min", [
    (4, "deepfm_criteo", "table_wise", 0),        # cfg3's sharding (SURVEY §8c golden (2): W in {1,2,4,8} x {TW,RW,mixed})
    (4, "dlrm_criteo", "row_wise", 0),
    (8, "dlrm_criteo", "row_wise", 0),            # cfg2's world size: tables smaller than W leave ranks empty
    (4, "mmoe_taobao", "mixed", 250),             # cfg5: big tables row-wise, small ones table-wise, two task towers
    (8, "deepfm_criteo", "table_wise", 0),        # W = 8 x {TW, mixed}: completes the W in {1,2,4,8} x {TW,RW,mixed} grid
    (8, "dlrm_criteo", "mixed", 200),
])
def test_w_invariance_at_larger_world_sizes(world, name, sharding, rw_min):
    _run(world, name, sharding, rw_min_rows=rw_min)


@pytest.mark.parametrize("opt", ["adam", "partial_rowwise_adam"])
def test_sharded_adam_two_ranks(opt):
    # the Adam variants on the owners (tzk_fused_bwd_ex path): every shard keeps both moments and its own step counter;
    # W-invariance of the updated tables as for Adagrad
    _run(2, "dlrm_criteo", "row_wise", sparse_opt=opt)


def test_din_sequence_two_ranks():
    _run(2, "multi_tower_din_taobao", "mixed", rw_min_rows=250)


def test_make_plan_is_deterministic_and_balanced():
    from torcheasyrec_b200.distributed import make_plan
    from torcheasyrec_b200.embedding_modules import EmbeddingBagConfig
    from torcheasyrec_b200.example_configs import CRITEO_HASH_SIZES

    tabs = [EmbeddingBagConfig(num_embeddings=h, embedding_dim=16, name=f"t{i}") for i, h in enumerate(CRITEO_HASH_SIZES)]
    plan = make_plan(tabs, 8, "table_wise")
    owners = [plan[f"t{i}"].owner for i in range(26)]
    big = [owners[i] for i, h in enumerate(CRITEO_HASH_SIZES) if h == 40000000]
    assert len(set(big)) == 5                      # the five 40M-row tables land on five different ranks
    assert plan == make_plan(tabs, 8, "table_wise")
    rw = make_plan(tabs, 8, "row_wise")
    assert rw["t1"].block == 4883 and rw["t5"].block == 1   # 39060/8 -> 4883 ; 3 rows -> 1 (ranks 3..7 empty)
    mixed = make_plan(tabs, 8, "mixed", rw_min_rows=255877)
    assert mixed["t0"].kind == "row_wise" and mixed["t5"].kind == "table_wise"
    cons = make_plan(tabs, 8, "row_wise", {"t0": ["table_wise"]})
    assert cons["t0"].kind == "table_wise" and cons["t1"].kind == "row_wise"
