"""Host logic and index algebra of the peer-memory sparse step (torcheasyrec_b200/peer_exchange.py + csrc/tzk_peer.cu)
on a box without a GPU.

W ranks live in one process as threads; "symmetric memory" is a registry of per-rank tensors every rank can see and the
device barrier is a threading.Barrier.  The kernels run in two flavours:
  * "model":  tests/oracle_backend.py's loop-level restatement of what each CUDA kernel does (same owner rule, same
              wire / slot arithmetic);
  * "source": csrc/tzk_peer.cu ITSELF (gathers, the three bucketize kernels, gradient publish, dense all-reduce),
              compiled for the host with tests/native/cuda_cpu_shim.h (one std::thread per CUDA thread, real
              __syncthreads, emulated warp shuffles) and called through the same ctypes signatures and pointer tables
              as on the GPU.  The owner-side sort / update (tzk_bwd.cu: CUB + PTX loads) stays the model there.
Everything above the kernels — `PeerState` itself, the wire capacity, barrier sites, the call order — is the real code.
Checked against the UNSHARDED collection on the same ids: outputs bit-equal, updated tables equal to 1e-6 (the owners
apply the 1/W gradient scale, one more rounding than the unsharded twin)."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
os.environ.setdefault("TZK_PEER_MIRROR_ROWS", "200")   # these tiny collections: some tables mirrored, some read remotely

from oracle_backend import OracleKernels  # noqa: E402

from torcheasyrec_b200 import functional as Fn  # noqa: E402
from torcheasyrec_b200 import peer_exchange  # noqa: E402
from torcheasyrec_b200.distributed import TABLE_WISE, _DimGroup, make_plan  # noqa: E402
from torcheasyrec_b200.embedding_modules import (EmbeddingBagCollection, EmbeddingBagConfig, EmbeddingCollection,  # noqa: E402
                                                 EmbeddingConfig, PoolingType, SparseOptimizerSpec,
                                                 output_names_by_table)
from torcheasyrec_b200.kernels import OPT_ADAGRAD  # noqa: E402

P, I32, I64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64


class _SimSymm:
    def __init__(self, t, everyone):
        self.t, self.everyone = t, everyone      # everyone[r] -> rank r's tensor of the same allocation

    @property
    def ptrs(self):                              # the pointer table the kernels get (complete after the host barrier)
        W = len(self.everyone)
        return (ctypes.c_uint64 * W)(*[self.everyone[r].data_ptr() for r in range(W)])


@pytest.fixture(scope="module")
def host_compiled_peer_lib(tmp_path_factory):
    exp = os.path.join(HERE, "native")
    out = str(tmp_path_factory.mktemp("shim") / "libtzk_peer_cpu.so")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DTZK_CPU_SHIM", "-Wno-unknown-pragmas", "-I", exp, "-x",
                    "c++", os.path.join(os.path.dirname(HERE), "torcheasyrec_b200", "csrc", "tzk_peer.cu"), "-shared",
                    "-fPIC", "-o", out], check=True)
    L = ctypes.CDLL(out)
    L.tzk_peer_pooled_gather_fwd.argtypes = [P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, P, I64, P, P, P]
    L.tzk_peer_pooled_gather_fwd_sel.argtypes = [P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, P, I64, P, P, P, I32, P]
    L.tzk_peer_seq_gather_fwd.argtypes = [P, P, P, P, P, P, P, I32, I32, I32, I32, I64, P, P, P, P]
    L.tzk_peer_mirror_refresh.argtypes = [P, I32, P, P, P, P, I32, P, P]
    L.tzk_peer_bucketize_workspace_bytes.restype = ctypes.c_size_t
    L.tzk_peer_bucketize_workspace_bytes.argtypes = [I32, I32, I32]
    L.tzk_peer_bucketize.argtypes = [P, P, I32, I32, I32, P, P, P, P, I32, I64, P, P, P, P, ctypes.c_size_t, P]
    L.tzk_peer_publish_grad.argtypes = [P, I64, P, P, P, P, I32, I32, P, I64, P]
    L.tzk_peer_allreduce_mean.argtypes = [P, I32, I64, P, P]
    L.tzk_peer_push_grad.argtypes = [P, P, I64, P, P, P, P, P, I32, I32, I64, I32, I32, I32, P]
    return L


class SourceKernels(OracleKernels):
    """OracleKernels with the peer kernels of csrc/tzk_peer.cu executed from their host-compiled source."""

    name = "oracle+peer-source"

    def __init__(self, lib):
        super().__init__()
        self.L = lib

    @staticmethod
    def _lay(lay):
        i32 = lambda xs: torch.tensor(list(xs), dtype=torch.int32)
        return i32(lay.dim), i32(lay.col), i32(lay.pool)

    def peer_mirror_refresh(self, tables, W, seg_rank, seg_src, seg_dst, seg_n, mirror):
        rc = self.L.tzk_peer_mirror_refresh(tables.ptrs, W, seg_rank.data_ptr(), seg_src.data_ptr(), seg_dst.data_ptr(),
                                            seg_n.data_ptr(), seg_rank.numel(), mirror.data_ptr(), None)
        assert rc == 0, rc

    def peer_pooled_gather_fwd(self, tables, rf_w_off, feat_rows, feat_block, feat_owner, lay, ids, offsets, B, W,
                               out=None, mirror=None, feat_mirror_off=None, feat_sel=None):
        dim, col, pool = self._lay(lay)
        out = torch.full((B, lay.total_dim), float("nan")) if out is None else out
        if feat_sel is not None:
            rc = self.L.tzk_peer_pooled_gather_fwd_sel(
                tables.ptrs, rf_w_off.data_ptr(), feat_rows.data_ptr(), feat_block.data_ptr(), feat_owner.data_ptr(),
                dim.data_ptr(), col.data_ptr(), pool.data_ptr(), ids.data_ptr(), offsets.data_ptr(), lay.num_features, B, W,
                (lay.max_dim + 3) // 4 * 4, out.data_ptr(), lay.total_dim, None if mirror is None else mirror.data_ptr(),
                None if feat_mirror_off is None else feat_mirror_off.data_ptr(), feat_sel.data_ptr(), feat_sel.numel(),
                None)
            assert rc == 0, rc
            return out
        rc = self.L.tzk_peer_pooled_gather_fwd(tables.ptrs, rf_w_off.data_ptr(), feat_rows.data_ptr(),
                                               feat_block.data_ptr(), feat_owner.data_ptr(), dim.data_ptr(),
                                               col.data_ptr(), pool.data_ptr(), ids.data_ptr(), offsets.data_ptr(),
                                               lay.num_features, B, W, (lay.max_dim + 3) // 4 * 4, out.data_ptr(),
                                               lay.total_dim, None if mirror is None else mirror.data_ptr(),
                                               None if feat_mirror_off is None else feat_mirror_off.data_ptr(), None)
        assert rc == 0, rc
        return out

    def peer_seq_gather_fwd(self, tables, rf_w_off, feat_rows, feat_block, feat_owner, lay, ids, offsets, B, W,
                            mirror=None, feat_mirror_off=None):
        D, nnz = lay.dim[0], ids.numel()
        out = torch.full((nnz, D), float("nan"))
        rc = self.L.tzk_peer_seq_gather_fwd(tables.ptrs, rf_w_off.data_ptr(), feat_rows.data_ptr(), feat_block.data_ptr(),
                                            feat_owner.data_ptr(), ids.data_ptr(), offsets.data_ptr(), lay.num_features,
                                            B, W, D, nnz, out.data_ptr(), None if mirror is None else mirror.data_ptr(),
                                            None if feat_mirror_off is None else feat_mirror_off.data_ptr(), None)
        assert rc == 0, rc
        return out

    def peer_bucketize(self, ids, offsets, F, B, W, feat_block, feat_owner, feat_rows, rf_key_base, pooled, cap,
                       wire_key, wire_idx, counts):
        nb = self.L.tzk_peer_bucketize_workspace_bytes(F, B, W)
        ws = torch.zeros(nb // 4 + 1, dtype=torch.int32)
        rc = self.L.tzk_peer_bucketize(ids.data_ptr(), offsets.data_ptr(), F, B, W, feat_block.data_ptr(),
                                       feat_owner.data_ptr(), feat_rows.data_ptr(), rf_key_base.data_ptr(), int(pooled),
                                       cap, wire_key.data_ptr(), wire_idx.data_ptr(), counts.data_ptr(), ws.data_ptr(),
                                       nb, None)
        assert rc == 0, rc

    def peer_publish_grad(self, grad, lay, offsets, B, dst):
        dim, col, pool = self._lay(lay)
        grad = grad.contiguous()
        rc = self.L.tzk_peer_publish_grad(grad.data_ptr(), grad.shape[1], col.data_ptr(), dim.data_ptr(), pool.data_ptr(),
                                          offsets.data_ptr(), lay.num_features, B, dst.data_ptr(), dst.shape[1], None)
        assert rc == 0, rc

    def peer_allreduce_mean(self, srcs, W, n, out):
        rc = self.L.tzk_peer_allreduce_mean(srcs.ptrs, W, n, out.data_ptr(), None)
        assert rc == 0, rc

    def peer_push_grad(self, recv, grad, lay, offsets, wire_idx, counts, me, W, cap, B, pooled):
        _, col, pool = self._lay(lay)
        grad = grad.contiguous()
        rc = self.L.tzk_peer_push_grad(recv.ptrs, grad.data_ptr(), grad.shape[1], col.data_ptr(), pool.data_ptr(),
                                       offsets.data_ptr(), wire_idx.data_ptr(), counts.data_ptr(), me, W, cap, B,
                                       lay.dim[0], int(pooled), None)
        assert rc == 0, rc


def _sim_mixin(registry, tbar, tag, device="cpu"):
    """Process / device plumbing of PeerBase replaced by an in-process model.  device="cuda": the W ranks share ONE GPU
    (tests/test_peer_gpu.py) — real kernels, real streams, pointer tables into the same device; the barrier drains the
    device before the threads meet."""

    class Sim:
        def _alloc(self, numel, dtype):
            n = getattr(self, "_n_alloc", 0)
            self._n_alloc = n + 1
            slot = registry.setdefault((tag, n), {})
            slot[self.me] = torch.zeros(max(int(numel), 1), dtype=dtype, device=device)
            return _SimSymm(slot[self.me], slot)

        def _host_barrier(self):
            if device != "cpu":
                torch.cuda.synchronize()
            tbar.wait()

        def _barrier(self, site):
            if device != "cpu":
                torch.cuda.synchronize()
            tbar.wait()

    return Sim


def _run_ranks(W, body):
    errors = []
    tbar = threading.Barrier(W)

    def main(r):
        try:
            body(r, tbar)
        except Exception:                       # a dead thread would leave the others in the barrier forever
            import traceback

            errors.append((r, traceback.format_exc()))
            tbar.abort()

    threads = [threading.Thread(target=main, args=(r,)) for r in range(W)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, "\n".join(f"rank {r}: {tb}" for r, tb in errors)


def _pooled_configs():
    mk = lambda n, rows, feats, pool=PoolingType.SUM: EmbeddingBagConfig(
        num_embeddings=rows, embedding_dim=16, name=n, feature_names=feats, pooling=pool)
    return [mk("t_big", 997, ["a"]), mk("t_shared", 64, ["b", "c"]), mk("t_tiny", 2, ["d"]),
            mk("t_mean", 301, ["e"], PoolingType.MEAN), mk("t_tw", 150, ["f"])]


def _bags(rng, F, B, feat_rows, multi_hot):
    """KJT ids / offsets of one rank: one id per bag, or ragged bags (0..4 ids, empty ones included)."""
    if not multi_hot:
        ids = np.concatenate([rng.integers(0, feat_rows[f], B) for f in range(F)]).astype(np.int64)
        return torch.from_numpy(ids), torch.arange(F * B + 1, dtype=torch.int64)
    lens = rng.integers(0, 5, F * B)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ids = np.concatenate([rng.integers(0, feat_rows[b // B], lens[b]) for b in range(F * B)] + [np.zeros(0, np.int64)])
    return torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(off)


def _cat_key_major(ids, offs, F, B, W):
    """Key-major concatenation of W per-rank KJTs (what the unsharded twin sees as one batch of W * B samples)."""
    out_ids, lens = [], []
    for f in range(F):
        for r in range(W):
            o = offs[r].numpy()
            out_ids.append(ids[r].numpy()[o[f * B]:o[(f + 1) * B]])
            lens.append(np.diff(o[f * B:(f + 1) * B + 1]))
    lens = np.concatenate(lens)
    return (torch.from_numpy(np.concatenate(out_ids).astype(np.int64)),
            torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)))


@pytest.mark.parametrize("kernels", ["model", "source"])
@pytest.mark.parametrize("W,multi_hot,tiny_rw", [(2, False, False), (3, False, False), (4, False, False), (2, True, False),
                                                 (3, True, False), (4, False, True), (8, False, True)])
def test_peer_step_matches_unsharded(W, multi_hot, tiny_rw, kernels, host_compiled_peer_lib):
    """tiny_rw: the 2-row table is split row-wise too (a table smaller than W leaves ranks without rows;
    997 rows over 8 ranks: hash_size % W != 0, short last shard)."""
    torch.manual_seed(0)
    rng = np.random.default_rng(7 + W)
    cfgs = _pooled_configs()
    B, D = 12, 16
    plan = make_plan(cfgs, W, "row_wise", {"t_tw": [TABLE_WISE]} if tiny_rw else {"t_tw": [TABLE_WISE], "t_tiny": [TABLE_WISE]})
    names = output_names_by_table(cfgs)
    spec = SparseOptimizerSpec(kind=OPT_ADAGRAD, lr=0.05)
    backend = OracleKernels() if kernels == "model" else SourceKernels(host_compiled_peer_lib)
    with Fn.use_backend(backend):
        full = EmbeddingBagCollection(cfgs, device="cpu")
        full.set_optimizer(spec)
        F = len(full.feature_names())
        feat_rows = [cfgs[t].num_embeddings for t in full._feat_table]
        batches = [_bags(rng, F, B, feat_rows, multi_hot) for _ in range(W)]
        ids, offs = [b[0] for b in batches], [b[1] for b in batches]
        grads = [torch.from_numpy(rng.standard_normal((B, F * D)).astype(np.float32)) for _ in range(W)]

        groups = []
        for r in range(W):
            g = _DimGroup(cfgs, plan, r, W, torch.device("cpu"), True, names)
            g.static_alpha = 2.5
            g.local.set_optimizer(spec)
            for t, c in enumerate(cfgs):
                n = g.local._table_rows[t]
                if n:
                    start = 0 if plan[c.name].kind == TABLE_WISE else r * plan[c.name].block
                    g.local.set_table_weight(t, full.table_weight(t)[start:start + n])
            groups.append(g)

        registry, outs = {}, [None] * W
        budget = [B * (4 if multi_hot else 1)] * F

        def body(r, tbar):
            class St(_sim_mixin(registry, tbar, "sparse"), peer_exchange.PeerState):
                pass

            st = St(groups[r], plan, None, B, budget)
            for _ in range(2):                    # two steps: wire buffers, counts and flags are reused
                outs[r] = st.gather(ids[r], offs[r])
                st.prep(ids[r], offs[r])
                st.backward(grads[r], offs[r])

        _run_ranks(W, body)
        assert all(int(g.overflow.item()) == 0 for g in groups)

        # forward (of the second step): same bits as the unsharded gather on that rank's ids after ONE update
        k = OracleKernels()
        cat_ids, cat_off = _cat_key_major(ids, offs, F, B, W)
        cat_grad = torch.cat(grads) / W
        k.fused_bwd(spec.kind, True, cat_grad, full.weights.data, full.opt_state, full.layout, cat_ids, cat_off, B * W,
                    spec.lr, spec.eps, 1.0)
        for r in range(W):
            want = k.pooled_gather_fwd(full.weights.data, full.layout, ids[r], offs[r], B)
            np.testing.assert_allclose(outs[r].numpy(), want.numpy(), rtol=2e-6, atol=1e-7)
        # second update on the twin, then the tables
        k.fused_bwd(spec.kind, True, cat_grad, full.weights.data, full.opt_state, full.layout, cat_ids, cat_off, B * W,
                    spec.lr, spec.eps, 1.0)
        for t, c in enumerate(cfgs):
            sh = plan[c.name]
            got = torch.zeros_like(full.table_weight(t))
            for r in range(W):
                n = groups[r].local._table_rows[t]
                if n:
                    start = 0 if sh.kind == TABLE_WISE else r * sh.block
                    got[start:start + n] = groups[r].local.table_weight(t)
            np.testing.assert_allclose(got.numpy(), full.table_weight(t).numpy(), rtol=2e-6, atol=1e-7, err_msg=c.name)


@pytest.mark.parametrize("kernels", ["model", "source"])
def test_peer_forward_is_bit_identical_to_the_unsharded_gather(kernels, host_compiled_peer_lib):
    """No update in between: every rank's pooled output equals the unsharded gather's bits (SUM and MEAN, ragged bags)."""
    W, B = 3, 9
    rng = np.random.default_rng(3)
    cfgs = _pooled_configs()
    plan = make_plan(cfgs, W, "row_wise", {"t_tw": [TABLE_WISE]})
    names = output_names_by_table(cfgs)
    backend = OracleKernels() if kernels == "model" else SourceKernels(host_compiled_peer_lib)
    with Fn.use_backend(backend):
        full = EmbeddingBagCollection(cfgs, device="cpu")
        F = len(full.feature_names())
        feat_rows = [cfgs[t].num_embeddings for t in full._feat_table]
        groups = []
        for r in range(W):
            g = _DimGroup(cfgs, plan, r, W, torch.device("cpu"), True, names)
            g.static_alpha = 2.0
            for t, c in enumerate(cfgs):
                n = g.local._table_rows[t]
                if n:
                    start = 0 if plan[c.name].kind == TABLE_WISE else r * plan[c.name].block
                    g.local.set_table_weight(t, full.table_weight(t)[start:start + n])
            groups.append(g)
        batches = [_bags(rng, F, B, feat_rows, True) for _ in range(W)]
        registry, outs = {}, [None] * W

        def body(r, tbar):
            class St(_sim_mixin(registry, tbar, "fwd"), peer_exchange.PeerState):
                pass

            st = St(groups[r], plan, None, B, [B * 4] * F)
            outs[r] = st.gather(*batches[r])

        _run_ranks(W, body)
        for r in range(W):
            want = OracleKernels().pooled_gather_fwd(full.weights.data, full.layout, batches[r][0], batches[r][1], B)
            np.testing.assert_array_equal(outs[r].numpy(), want.numpy())


@pytest.mark.parametrize("kernels", ["model", "source"])
@pytest.mark.parametrize("W", [2, 3])
def test_peer_sequence_collection_matches_unsharded(W, kernels, host_compiled_peer_lib):
    """EmbeddingCollection (un-pooled, ragged sequences): rows bit-equal, tables after one update equal to 1e-6."""
    rng = np.random.default_rng(11 + W)
    mk = lambda n, rows, feats: EmbeddingConfig(num_embeddings=rows, embedding_dim=8, name=n, feature_names=feats)
    cfgs = [mk("q", 50, ["q_id"]), mk("s1", 211, ["seq_a"]), mk("s2", 40, ["seq_b"])]
    B, D = 7, 8
    plan = make_plan(cfgs, W, "row_wise", {"s2": [TABLE_WISE]})
    names = output_names_by_table(cfgs)
    spec = SparseOptimizerSpec(kind=OPT_ADAGRAD, lr=0.1)
    backend = OracleKernels() if kernels == "model" else SourceKernels(host_compiled_peer_lib)
    with Fn.use_backend(backend):
        full = EmbeddingCollection(cfgs, device="cpu")
        full.set_optimizer(spec)
        F = len(full.feature_names())
        feat_rows = [cfgs[t].num_embeddings for t in full._feat_table]
        max_len = 6
        batches = []
        for _ in range(W):
            lens = np.concatenate([np.ones(B, np.int64), rng.integers(0, max_len + 1, B), rng.integers(0, max_len + 1, B)])
            off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            ids = np.concatenate([rng.integers(0, feat_rows[b // B], lens[b]) for b in range(F * B)]).astype(np.int64)
            batches.append((torch.from_numpy(ids), torch.from_numpy(off)))
        grads = [torch.from_numpy(rng.standard_normal((b[0].numel(), D)).astype(np.float32)) for b in batches]
        groups = []
        for r in range(W):
            g = _DimGroup(cfgs, plan, r, W, torch.device("cpu"), False, names)
            g.static_alpha = 2.0
            g.local.set_optimizer(spec)
            for t, c in enumerate(cfgs):
                n = g.local._table_rows[t]
                if n:
                    start = 0 if plan[c.name].kind == TABLE_WISE else r * plan[c.name].block
                    g.local.set_table_weight(t, full.table_weight(t)[start:start + n])
            groups.append(g)
        registry, outs = {}, [None] * W

        def body(r, tbar):
            class St(_sim_mixin(registry, tbar, "seq"), peer_exchange.PeerState):
                pass

            st = St(groups[r], plan, None, B, [B, B * max_len, B * max_len])
            outs[r] = st.gather(*batches[r])
            st.prep(*batches[r])
            st.backward(grads[r], batches[r][1])

        _run_ranks(W, body)
        assert all(int(g.overflow.item()) == 0 for g in groups)
        k = OracleKernels()
        for r in range(W):
            want = k.seq_gather_fwd(full.weights.data, full.layout, batches[r][0], batches[r][1], B)
            np.testing.assert_array_equal(outs[r].numpy(), want.numpy())
        # the unsharded twin: key-major concatenation of ids and of the gradient rows
        ids, offs = [b[0] for b in batches], [b[1] for b in batches]
        cat_ids, cat_off = _cat_key_major(ids, offs, F, B, W)
        rows = []
        for f in range(F):
            for r in range(W):
                o = offs[r].numpy()
                rows.append(grads[r][o[f * B]:o[(f + 1) * B]])
        k.fused_bwd(spec.kind, False, torch.cat(rows) / W, full.weights.data, full.opt_state, full.layout, cat_ids,
                    cat_off, B * W, spec.lr, spec.eps, 1.0)
        for t, c in enumerate(cfgs):
            sh = plan[c.name]
            got = torch.zeros_like(full.table_weight(t))
            for r in range(W):
                n = groups[r].local._table_rows[t]
                if n:
                    start = 0 if sh.kind == TABLE_WISE else r * sh.block
                    got[start:start + n] = groups[r].local.table_weight(t)
            np.testing.assert_allclose(got.numpy(), full.table_weight(t).numpy(), rtol=2e-6, atol=1e-7, err_msg=c.name)


@pytest.mark.parametrize("kernels", ["model", "source"])
def test_peer_overflow_is_reported_on_every_rank(kernels, host_compiled_peer_lib):
    """All ids of every rank hit rank 0's block: with a tight wire capacity ids are dropped, nothing is corrupted
    (updates stay inside the right tables) and EVERY rank's overflow flag is raised in the same step."""
    W, B = 3, 16
    cfgs = [EmbeddingBagConfig(num_embeddings=300, embedding_dim=16, name="t", feature_names=["a"]),
            EmbeddingBagConfig(num_embeddings=90, embedding_dim=16, name="u", feature_names=["b"])]
    plan = make_plan(cfgs, W, "row_wise")
    names = output_names_by_table(cfgs)
    spec = SparseOptimizerSpec(kind=OPT_ADAGRAD, lr=0.05)
    backend = OracleKernels() if kernels == "model" else SourceKernels(host_compiled_peer_lib)
    with Fn.use_backend(backend):
        groups = []
        for r in range(W):
            g = _DimGroup(cfgs, plan, r, W, torch.device("cpu"), True, names)
            g.static_alpha = 1.0
            g.local.set_optimizer(spec)
            groups.append(g)
        before = [g.local.weights.data.clone() for g in groups]
        ids = torch.cat([torch.arange(B) % 7, torch.arange(B) % 5]).to(torch.int64)      # all in rank 0's blocks
        off = torch.arange(2 * B + 1, dtype=torch.int64)
        registry = {}

        def body(r, tbar):
            class St(_sim_mixin(registry, tbar, "ovf"), peer_exchange.PeerState):
                pass

            st = St(groups[r], plan, None, B)
            st.gather(ids, off)
            st.prep(ids, off)
            st.backward(torch.ones(B, 32), off)

        _run_ranks(W, body)
        assert all(int(g.overflow.item()) == 1 for g in groups)
        for r in (1, 2):                      # nobody but rank 0 owns a touched row
            assert torch.equal(groups[r].local.weights.data, before[r])
        assert not torch.equal(groups[0].local.weights.data, before[0])


@pytest.mark.parametrize("kernels", ["model", "source"])
def test_peer_dense_grad_sync_is_the_rank_ordered_mean(kernels, host_compiled_peer_lib):
    W = 4
    backend = OracleKernels() if kernels == "model" else SourceKernels(host_compiled_peer_lib)
    rng = np.random.default_rng(5)
    vals = [[torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for s in ((3, 5), (7,), (2, 2))]
            for _ in range(W)]
    params = [[torch.nn.Parameter(torch.zeros(s)) for s in ((3, 5), (7,), (2, 2))] for _ in range(W)]
    registry, got = {}, [None] * W
    with Fn.use_backend(backend):
        def body(r, tbar):
            class Sync(_sim_mixin(registry, tbar, "dense"), peer_exchange.PeerDenseGradSync):
                pass

            s = Sync(params[r], None, world=W, rank=r)
            for step in range(2):
                s.zero()
                for p, v in zip(params[r], vals[r]):
                    p.grad.add_(v * (step + 1))
                s.sync()
            got[r] = [p.grad.clone() for p in params[r]]

        _run_ranks(W, body)
    for i in range(3):
        acc = vals[0][i].numpy() * np.float32(2)
        for r in range(1, W):
            acc = acc + vals[r][i].numpy() * np.float32(2)
        want = acc * np.float32(1.0 / W)
        for r in range(W):
            np.testing.assert_array_equal(got[r][i].numpy(), want)          # identical bits on every rank
