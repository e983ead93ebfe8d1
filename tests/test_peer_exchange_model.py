"""Host logic and index algebra of the peer-memory sparse step (scripts/experimental/peer_exchange.py, round-2
groundwork — its CUDA kernels in scripts/experimental/tzk_peer.cu have not run on hardware yet).

W ranks live in one process as threads; "symmetric memory" is a registry of per-rank tensors every rank can see and the
device barrier is a threading.Barrier.  The peer kernels run in two flavours:
  * "model":  a loop-level restatement of what each CUDA kernel does (same owner rule, same wire / slot arithmetic);
  * "source": tzk_peer.cu ITSELF, compiled for the host with scripts/experimental/cuda_cpu_shim.h (one std::thread per
              CUDA thread, real __syncthreads) and called through the same ctypes signatures and pointer tables as on
              the GPU — index arithmetic, guards and argument marshalling of the real source, minus PTX and timing.
Everything above the kernels — `PeerState` itself, the wire capacity, `bounds`, `owner_layout_static`, the call into
the fused update — is the real code.  Checked against the
UNSHARDED collection on the same ids: pooled outputs bit-equal, updated tables equal to 1e-6 (the owners apply the
1/W gradient scale, one more rounding than the unsharded twin)."""
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
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "scripts", "experimental"))

from oracle_backend import OracleKernels  # noqa: E402

from torcheasyrec_b200 import functional as Fn  # noqa: E402
from torcheasyrec_b200.distributed import TABLE_WISE, _DimGroup, make_plan  # noqa: E402
from torcheasyrec_b200.embedding_modules import (EmbeddingBagCollection, EmbeddingBagConfig, PoolingType,  # noqa: E402
                                                 SparseOptimizerSpec, output_names_by_table)
from torcheasyrec_b200.kernels import OPT_ADAGRAD  # noqa: E402

peer_exchange = pytest.importorskip("peer_exchange")


class _SimSymm:
    def __init__(self, t, everyone):
        self.t, self.everyone = t, everyone      # everyone[r] -> rank r's tensor of the same allocation

    @property
    def ptrs(self):                              # the pointer table the kernels get (complete after the host barrier)
        W = len(self.everyone)
        return (ctypes.c_uint64 * W)(*[self.everyone[r].data_ptr() for r in range(W)])


@pytest.fixture(scope="module")
def host_compiled_peer_lib(tmp_path_factory):
    exp = os.path.join(os.path.dirname(HERE), "scripts", "experimental")
    out = str(tmp_path_factory.mktemp("shim") / "libtzk_peer_cpu.so")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-DTZK_CPU_SHIM", "-Wno-unknown-pragmas", "-x", "c++",
                    os.path.join(exp, "tzk_peer.cu"), "-shared", "-fPIC", "-o", out], check=True)
    return peer_exchange.declare(ctypes.CDLL(out))


def _make_source_state(registry, tbar, lib):
    """PeerState with its own _k_gather / _k_pull_counts / _k_pull (ctypes calls into the host-compiled CUDA source)."""
    Model = _make_sim_state(registry, tbar)

    class SourcePeerState(Model):
        def _init_io(self):
            super()._init_io()
            self.lib = lib

        _k_gather = peer_exchange.PeerState._k_gather
        _k_pull_counts = peer_exchange.PeerState._k_pull_counts
        _k_pull = peer_exchange.PeerState._k_pull

    return SourcePeerState


def _make_sim_state(registry, tbar):
    class SimPeerState(peer_exchange.PeerState):
        def _init_io(self):
            self._n_alloc = 0

        def _alloc(self, numel, dtype):
            slot = registry.setdefault(self._n_alloc, {})
            self._n_alloc += 1
            slot[self.me] = torch.zeros(max(int(numel), 1), dtype=dtype)
            return _SimSymm(slot[self.me], slot)

        def _host_barrier(self):
            tbar.wait()

        def _k_barrier(self):
            tbar.wait()

        # peer_pooled_gather_fwd_kernel: owner = feat_owner + id / block, clamped to the last rank; local row = id - q*block
        def _k_gather(self, ids, offsets, out):
            g, lay = self.g, self.g.local.layout
            F, B, W = g.F, self.B, self.W
            blocks, owners = g.feat_block.tolist(), g.feat_owner.tolist()
            rows, w_off = self.feat_rows.tolist(), self.rf_w_off.tolist()
            idl, off = ids.tolist(), offsets.tolist()
            o = out.numpy()
            for f in range(F):
                D, col = lay.dim[f], lay.col[f]
                for b in range(B):
                    s, e = off[f * B + b], off[f * B + b + 1]
                    acc = np.zeros(D, dtype=np.float32)
                    for l in range(s, e):
                        i = idl[l] if 0 <= idl[l] < rows[f] else 0
                        q = i // blocks[f]
                        r = owners[f] + q
                        if r >= W:
                            q -= r - (W - 1)
                            r = W - 1
                        base = w_off[r * F + f] + (i - q * blocks[f]) * D
                        acc = acc + self.tables.everyone[r].numpy()[base:base + D]
                    if lay.pool[f] == 1 and e > s:
                        acc = acc * np.float32(1.0 / (e - s))
                    o[b, col:col + D] = acc

        def _k_bucketize(self, ids, offsets):
            g = self.g
            _, oo, oi, op, _ = Fn.backend().bucketize_rw(ids, offsets, g.F, self.B, self.W, g.feat_block, want_pos=True,
                                                         feat_owner=g.feat_owner, wire_capacity=self.cap)
            self.wire_ids.t.copy_(oi)
            self.wire_pos.t.copy_(op)
            return oo

        def _k_pull_counts(self, recv_counts):
            F = self.g.F
            for r in range(self.W):
                recv_counts[r] = self.counts.everyone[r][self.me * F:(self.me + 1) * F]

        # peer_pull_kernel: slot s = (src r, j); valid while s is before r's padding run
        def _k_pull(self, bounds, recv_ids, recv_g):
            g = self.g
            F, B, D, cap = g.F, self.B, g.dim, self.cap
            col = g.local.layout.col
            bnd = bounds.tolist()
            for s in range(self.W * cap):
                r, j = divmod(s, cap)
                if s < bnd[r * (F + 1) + F]:
                    recv_ids[s] = self.wire_ids.everyone[r][self.me * cap + j]
                    pos = int(self.wire_pos.everyone[r][self.me * cap + j])
                    f, b = divmod(pos, B)
                    recv_g[s] = self.grad.everyone[r].view(B, g.total_dim)[b, col[f]:col[f] + D]
                else:
                    recv_ids[s] = 0
                    recv_g[s] = 0.0

    return SimPeerState


def _configs():
    mk = lambda n, rows, feats, pool=PoolingType.SUM: EmbeddingBagConfig(
        num_embeddings=rows, embedding_dim=16, name=n, feature_names=feats, pooling=pool)
    return [mk("t_big", 997, ["a"]), mk("t_shared", 64, ["b", "c"]), mk("t_tiny", 2, ["d"]),
            mk("t_mean", 301, ["e"], PoolingType.MEAN), mk("t_tw", 150, ["f"])]


@pytest.mark.parametrize("kernels", ["model", "source"])
@pytest.mark.parametrize("W", [2, 3, 4])
def test_peer_step_matches_unsharded(W, kernels, host_compiled_peer_lib):
    torch.manual_seed(0)
    rng = np.random.default_rng(7)
    cfgs = _configs()
    B, D = 12, 16
    plan = make_plan(cfgs, W, "row_wise", {"t_tw": [TABLE_WISE], "t_tiny": [TABLE_WISE]})
    names = output_names_by_table(cfgs)
    spec = SparseOptimizerSpec(kind=OPT_ADAGRAD, lr=0.05)
    with Fn.use_backend(OracleKernels()):
        full = EmbeddingBagCollection(cfgs, device="cpu")
        full.set_optimizer(spec)
        F = len(full.feature_names())
        feat_rows = [cfgs[t].num_embeddings for t in full._feat_table]
        # one id per bag, per rank its own batch
        ids = [torch.from_numpy(np.concatenate([rng.integers(0, feat_rows[f], B) for f in range(F)]).astype(np.int64))
               for _ in range(W)]
        offsets = torch.arange(F * B + 1, dtype=torch.int64)
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

        registry, tbar = {}, threading.Barrier(W)
        Sim = (_make_sim_state(registry, tbar) if kernels == "model"
               else _make_source_state(registry, tbar, host_compiled_peer_lib))
        states, outs, errors = [None] * W, [None] * W, []

        def rank_main(r):
            try:
                st = Sim(groups[r], plan, None, B)
                states[r] = st
                outs[r] = st.gather(ids[r], offsets)
                st.bucketize(ids[r], offsets)
                st.backward(grads[r])
            except Exception:                       # a dead thread would leave the others in the barrier forever
                import traceback

                errors.append((r, traceback.format_exc()))
                tbar.abort()

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(W)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=120)
        assert not errors, "\n".join(f"rank {r}: {tb}" for r, tb in errors)
        assert all(int(g.overflow.item()) == 0 for g in groups)

        # forward: same bits as the unsharded gather on that rank's ids
        k = Fn.backend()
        for r in range(W):
            want = k.pooled_gather_fwd(full.weights.data, full.layout, ids[r], offsets, B)
            np.testing.assert_array_equal(outs[r].numpy(), want.numpy())

        # backward: the unsharded twin on the concatenated batch (key-major concat), gradient / W
        cat_ids = torch.cat([torch.cat([ids[r][f * B:(f + 1) * B] for r in range(W)]) for f in range(F)])
        cat_off = torch.arange(F * B * W + 1, dtype=torch.int64)
        cat_grad = torch.cat(grads) / W
        k.fused_bwd(spec.kind, True, cat_grad, full.weights.data, full.opt_state, full.layout, cat_ids, cat_off, B * W,
                    spec.lr, spec.eps, 1.0)
        for t, c in enumerate(cfgs):
            sh = plan[c.name]
            parts = []
            for r in range(W):
                n = groups[r].local._table_rows[t]
                if n:
                    parts.append((0 if sh.kind == TABLE_WISE else r * sh.block, groups[r].local.table_weight(t)))
            got = torch.zeros_like(full.table_weight(t))
            for start, w in parts:
                got[start:start + w.shape[0]] = w
            np.testing.assert_allclose(got.numpy(), full.table_weight(t).numpy(), rtol=1e-6, atol=1e-7, err_msg=c.name)
