"""Host side of csrc/tzk_peer.cu: the sharded sparse step over peer memory (exchange="peer" in shard_model).

`enable_peer_exchange(sharded_ebc)` re-routes a `ShardedEmbeddingBagCollection` that was built with a static wire
capacity (`shard_model(..., static_capacity=a)`, one id per bag) from the three NCCL all-to-alls per dim group to
peer-memory kernels over the NVSwitch domain:

    forward    one kernel: the requester gathers rows straight out of the owners' arenas (symmetric memory) and pools.
    backward   bucketize into the rank's own symmetric wire buffer (runs during the forward pass), publish the
               pooled-output gradient, barrier, the owner pulls ids + gradient slices, tzk_fused_bwd as before, barrier.

Symmetric allocations and the address exchange come from `torch.distributed._symmetric_memory` (plumbing); every
kernel on the path is ours.  The whole step stays capturable: the barrier's epoch lives on the device.

What is the same as the NCCL static path (so results are bit-identical to it): the wire layout (destination-major,
feature runs, fixed capacity), the owner-side `bounds` / `owner_layout_static`, the update kernel and its 1/W scale.
"""
import ctypes
from typing import List

import torch
import torch.distributed as dist

from . import functional as Fn
from .kernels import build_layout
from .sparse import KeyedTensor

def load_lib():
    """The package library (tzk_peer_* entry points are part of libtzk.so)."""
    from ._lib import lib

    return lib()


def _check(rc: int, what: str) -> None:
    if rc:
        raise RuntimeError(f"{what} failed with code {rc}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0


class _Symm:
    """One symmetric allocation: `t` is this rank's tensor, `ptrs` a host array with every rank's address as mapped in
    this process (rank order).  Allocation + address exchange: `torch.distributed._symmetric_memory` (cuMem VMM handles);
    TZK_PEER_ALLOC=ipc — or a failing rendezvous — falls back to classic CUDA IPC handles through
    torch.multiprocessing's reductions (cudaIpcGetMemHandle / cudaIpcOpenMemHandle), one all_gather_object each."""

    def __init__(self, numel: int, dtype, device, group) -> None:
        import os

        W = dist.get_world_size(group)
        n = max(int(numel), 1)
        self.h = None
        if os.environ.get("TZK_PEER_ALLOC", "symm") != "ipc":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                self.t = symm_mem.empty(n, dtype=dtype, device=device)
                self.t.zero_()
                self.h = symm_mem.rendezvous(self.t, group)
                self.ptrs = (ctypes.c_uint64 * W)(*[int(p) for p in self.h.buffer_ptrs])
                return
            except Exception as e:  # noqa: BLE001 — e.g. no pidfd / fabric handle support in this container
                import warnings

                warnings.warn(f"symmetric-memory rendezvous failed ({e!r}); falling back to CUDA IPC handles")
        self._ipc(n, dtype, device, group, W)

    def _ipc(self, n: int, dtype, device, group, W: int) -> None:
        from torch.multiprocessing.reductions import reduce_tensor

        me = dist.get_rank(group)
        # a private cudaMalloc block (IPC shares whole allocations): ask the caching allocator for an exclusive segment
        self.t = torch.zeros(n, dtype=dtype, device=device)
        torch.cuda.synchronize()
        fn, args = reduce_tensor(self.t)
        gathered = [None] * W
        dist.all_gather_object(gathered, args, group=group)
        self.peers = [self.t if r == me else fn(*gathered[r]) for r in range(W)]
        for r, pt in enumerate(self.peers):     # first touch enables peer access between the two devices
            if r != me:
                _ = pt[:1].to(device)
        torch.cuda.synchronize()
        self.ptrs = (ctypes.c_uint64 * W)(*[int(pt.data_ptr()) for pt in self.peers])


class PeerState:
    """Peer-memory state of one `_DimGroup` (all tables of one embedding dim)."""

    def __init__(self, g, plan, group, batch_size: int) -> None:
        if not g.static_alpha:
            raise ValueError("peer exchange needs the static wire capacity (shard_model(..., static_capacity=a))")
        self.g, self.group, self.B = g, group, int(batch_size)
        self.W, self.me = g.world, g.rank
        dev, F, W = g.device, g.F, g.world
        self._init_io()
        lay = g.local.layout
        # every rank's arena layout (deterministic from the plan: no communication)
        from .distributed import local_rows

        per_rank = [build_layout([local_rows(c, plan[c.name], r) for c in g.configs], g.local._table_dim,
                                 g.local._feat_table, list(lay.pool)) for r in range(W)]
        assert per_rank[self.me].w_off == list(lay.w_off), "local layout differs from the plan's"
        self.rf_w_off = torch.tensor([o for lr in per_rank for o in lr.w_off], dtype=torch.int64, device=dev)
        self.feat_rows = torch.tensor([g.configs[t].num_embeddings for t in g.local._feat_table], dtype=torch.int64,
                                      device=dev)
        # the arena moves into symmetric memory (same size on every rank: the largest shard)
        self.tables = self._alloc(max(lr.arena_elems for lr in per_rank), torch.float32)
        n = g.local.weights.numel()
        self.tables.t[:n].copy_(g.local.weights.data)
        g.local.weights.data = self.tables.t[:n]
        # wire buffers (sized on first use: nnz = F * B for one id per bag), gradient, flags
        nnz = F * self.B
        self.cap = (int(g.static_alpha * nnz / W) + 8) // 8 * 8
        g.static_nnz, g.static_cap = nnz, self.cap
        self.wire_ids = self._alloc(W * self.cap, torch.int64)
        self.wire_pos = self._alloc(W * self.cap, torch.int32)
        self.counts = self._alloc(W * F, torch.int32)
        self.grad = self._alloc(self.B * g.total_dim, torch.float32)
        self.pads = self._alloc(W, torch.int32)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self._host_barrier()                 # flags are zero and tables are in place everywhere before the first step

    # ---- device / process plumbing (tests/test_peer_exchange_model.py swaps these for an in-process model) ----------
    def _init_io(self) -> None:
        self.lib = load_lib()

    def _alloc(self, numel: int, dtype) -> "_Symm":
        return _Symm(numel, dtype, self.g.device, self.group)

    def _host_barrier(self) -> None:
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def _k_barrier(self) -> None:
        _check(self.lib.tzk_peer_barrier(self.pads.ptrs, self.me, self.W, self.epoch.data_ptr(), _stream()), "barrier")

    def _k_gather(self, ids, offsets, out) -> None:
        g, lay = self.g, self.g.local.layout
        _check(self.lib.tzk_peer_pooled_gather_fwd(
            self.tables.ptrs, self.rf_w_off.data_ptr(), self.feat_rows.data_ptr(), g.feat_block.data_ptr(),
            g.feat_owner.data_ptr(), lay.d_dim.data_ptr(), lay.d_col.data_ptr(), lay.d_pool.data_ptr(), ids.data_ptr(),
            offsets.data_ptr(), g.F, self.B, self.W, lay.max_dim, out.data_ptr(), g.total_dim, _stream()),
            "peer_pooled_gather_fwd")

    def _k_bucketize(self, ids, offsets) -> torch.Tensor:
        """tzk_bucketize_rw straight into this rank's wire buffers; returns out_offsets [W*F*B+1]."""
        g, k = self.g, Fn.backend()
        F, B, W, nnz = g.F, self.B, self.W, ids.numel()
        out_lengths = torch.empty(W * F * B, dtype=torch.int32, device=ids.device)
        oo = torch.empty(W * F * B + 1, dtype=torch.int64, device=ids.device)
        ws = k._workspace("bucketize", k._lib.tzk_bucketize_rw_workspace_bytes(F, B, W, nnz), ids.device)
        _check(k._lib.tzk_bucketize_rw(ids.data_ptr(), offsets.data_ptr(), F, B, W, g.feat_block.data_ptr(),
                                       g.feat_owner.data_ptr(), nnz, self.cap, out_lengths.data_ptr(), oo.data_ptr(),
                                       self.wire_ids.t.data_ptr(), self.wire_pos.t.data_ptr(), None, ws.data_ptr(),
                                       ws.numel(), _stream()), "tzk_bucketize_rw")
        return oo

    def _k_pull_counts(self, recv_counts) -> None:
        _check(self.lib.tzk_peer_pull_counts(self.counts.ptrs, self.me, self.W, self.g.F, recv_counts.data_ptr(),
                                             _stream()), "peer_pull_counts")

    def _k_pull(self, bounds, recv_ids, recv_g) -> None:
        g = self.g
        _check(self.lib.tzk_peer_pull(self.wire_ids.ptrs, self.wire_pos.ptrs, self.grad.ptrs, self.me, self.W, self.cap,
                                      g.F, self.B, g.dim, g.local.layout.d_col.data_ptr(), bounds.data_ptr(), g.total_dim,
                                      recv_ids.data_ptr(), recv_g.data_ptr(), _stream()), "peer_pull")

    # ---- the step ------------------------------------------------------------------------------------------------
    def barrier(self) -> None:
        self._k_barrier()

    def gather(self, ids: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.B, self.g.total_dim), dtype=torch.float32, device=ids.device)
        self._k_gather(ids, offsets, out)
        return out

    def bucketize(self, ids: torch.Tensor, offsets: torch.Tensor) -> None:
        """ids -> this rank's own wire buffers (destination r at slot r * cap) + counts[dest, f]."""
        g = self.g
        F, B = g.F, self.B
        if ids.numel() != F * B:
            raise RuntimeError(f"peer exchange is sized for one id per bag ({F * B} ids), got {ids.numel()}")
        oo = self._k_bucketize(ids, offsets)
        seg = oo[::B]
        self.counts.t.copy_((seg[1:] - seg[:-1]).to(torch.int32))
        dest_start = oo[::F * B]
        g.overflow.add_(((dest_start[1:] - dest_start[:-1]) > self.cap).any().to(torch.int32))

    def backward(self, grad_out: torch.Tensor) -> None:
        g, k = self.g, Fn.backend()
        F, W, cap, D = g.F, self.W, self.cap, g.dim
        spec = g.local.optimizer
        if spec is None:
            raise RuntimeError("sharded collection: no sparse optimizer set (call set_optimizer)")
        self.grad.t.view(self.B, g.total_dim).copy_(grad_out)
        self.barrier()                                   # every rank's wire buffers and gradient are published
        recv_counts = torch.empty((W, F), dtype=torch.int32, device=grad_out.device)
        self._k_pull_counts(recv_counts)
        tot = recv_counts.sum(1, keepdim=True)
        lens = torch.cat([recv_counts, (cap - tot).clamp_(min=0)], dim=1).reshape(-1).to(torch.int32)
        bounds = k.lengths_to_offsets(lens)              # [W * (F + 1) + 1], as in _StaticDispatch
        recv_ids = torch.empty(W * cap, dtype=torch.int64, device=grad_out.device)
        recv_g = torch.empty((W * cap, D), dtype=torch.float32, device=grad_out.device)
        self._k_pull(bounds, recv_ids, recv_g)
        k.fused_bwd(spec.kind, False, recv_g, g.local.weights.data, g.local.opt_state, g.owner_layout_static, recv_ids,
                    bounds, 1, spec.lr, spec.eps, 1.0 / W, **g.local.opt_extras())
        self.barrier()                                   # tables quiescent, wire buffers / gradient reusable


class _PeerPooled(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hook, st: PeerState, ids, offsets):
        out = st.gather(ids, offsets)
        if hook is not None:
            st.bucketize(ids, offsets)      # only the backward needs it; TODO(side stream, like the early sort)
        ctx.st = st
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ctx.st.backward(grad_out.contiguous())
        return None, None, None, None


def enable_peer_exchange(sm, batch_size: int) -> List[PeerState]:
    """Switches `sm` (static-capacity sharded pooled collection) to the peer-memory path for local batches of
    `batch_size` samples.  Returns the per-dim-group states (kept alive by the patched forward)."""
    states = [PeerState(g, sm.plan, sm._group if sm._group is not None else dist.group.WORLD, batch_size)
              for g in sm.groups]

    def forward(features):
        keys, lens, vals = [], [], []
        for g, st in zip(sm.groups, states):
            kjt = g.local._select(features)
            if kjt.stride() != st.B:
                raise RuntimeError(f"peer exchange was sized for batch {st.B}, got {kjt.stride()}")
            vals.append(_PeerPooled.apply(sm._hook_tensor(kjt.values().device), st, kjt.values(), kjt.offsets()))
            keys += g.embedding_names
            lens += [g.dim] * g.F
        return KeyedTensor(keys, lens, vals[0] if len(vals) == 1 else torch.cat(vals, dim=1))

    sm.forward = forward
    sm._peer_states = states
    return states
