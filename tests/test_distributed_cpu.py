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
        from torcheasyrec_b200.verify import verify_sharded

        # CPU: host logic over gloo with the oracle as compute; GPU: the CUDA kernels over NCCL / peer memory
        with (contextlib.nullcontext() if use_cuda else Fn.use_backend(OracleKernels())):
            verify_sharded(name, dev, sharding, rw_min_rows=rw_min_rows, static_capacity=static_capacity,
                           exchange=exchange, sparse_opt=sparse_opt, bit_exact_logits=not use_cuda)
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
