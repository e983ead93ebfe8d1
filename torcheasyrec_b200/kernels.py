"""Tensor-level wrappers over the tzk C-ABI (include/tzk.h).

`CudaKernels` is the only compute backend the package ships.  It validates device / dtype / contiguity,
allocates outputs and workspaces through torch's caching allocator, passes raw pointers + the current
CUDA stream across the ABI, and raises `TzkError` on any failure.  There is deliberately no CPU
implementation here: host-logic tests inject their own checker backend (tests/oracle_backend.py).
"""

import ctypes
import os
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from ._lib import TzkError, check, lib

POOL_SUM, POOL_MEAN = 0, 1
OPT_SGD, OPT_ADAGRAD, OPT_ROWWISE_ADAGRAD, OPT_ADAM, OPT_PARTIAL_ROWWISE_ADAM = 0, 1, 2, 3, 4
OPT_ACCUM_OUT = 100      # peer-memory step: per-row gradient sums into a dense buffer instead of an update


@dataclass
class FeatureLayout:
    """Host-side description of the keys served by one shard arena (one entry per KJT key).

    Mirrors the "feature descriptor" arrays of include/tzk.h.  `key_base` linearises (table,row) for the
    backward sort; features that share a physical table share w_off / key_base.
    """

    w_off: List[int]
    rows: List[int]
    dim: List[int]
    col: List[int]
    pool: List[int]
    key_base: List[int]
    total_keys: int
    total_dim: int
    arena_elems: int
    # elements between consecutive rows of each key's table: None = dense rows (= dim); `interleaved` = every table row is
    # followed by its element-wise optimizer-state row (stride 2 * dim; tzk_opt_args.interleaved, include/tzk.h)
    stride: Optional[List[int]] = None
    interleaved: bool = False
    d_stride: Optional[torch.Tensor] = None
    # device copies (filled by .to())
    d_w_off: Optional[torch.Tensor] = None
    d_rows: Optional[torch.Tensor] = None
    d_dim: Optional[torch.Tensor] = None
    d_col: Optional[torch.Tensor] = None
    d_pool: Optional[torch.Tensor] = None
    d_key_base: Optional[torch.Tensor] = None

    @property
    def num_features(self) -> int:
        return len(self.dim)

    @property
    def max_dim(self) -> int:
        return max(self.dim) if self.dim else 1

    @property
    def vec_ok(self) -> int:
        return int(all(d % 4 == 0 for d in self.dim) and all(c % 4 == 0 for c in self.col)
                   and all(o % 4 == 0 for o in self.w_off) and all(s % 4 == 0 for s in (self.stride or [])))

    def row_stride(self, f: int) -> int:
        return self.stride[f] if self.stride is not None else self.dim[f]

    def to(self, device) -> "FeatureLayout":
        self.d_w_off = torch.tensor(self.w_off, dtype=torch.int64, device=device)
        self.d_rows = torch.tensor(self.rows, dtype=torch.int64, device=device)
        self.d_dim = torch.tensor(self.dim, dtype=torch.int32, device=device)
        self.d_col = torch.tensor(self.col, dtype=torch.int32, device=device)
        self.d_pool = torch.tensor(self.pool, dtype=torch.int32, device=device)
        self.d_key_base = torch.tensor(self.key_base, dtype=torch.int64, device=device)
        self.d_stride = None if self.stride is None else torch.tensor(self.stride, dtype=torch.int32, device=device)
        return self


def build_layout(table_rows: Sequence[int], table_dim: Sequence[int], feat_table: Sequence[int],
                 feat_pool: Sequence[int], align: int = 4, interleaved: bool = False) -> FeatureLayout:
    """Packs tables back to back into one arena (row starts 16-B aligned) and lays features out in order.
    `interleaved`: every table row is followed by its optimizer-state row (row stride 2 * dim; table starts 128-B
    aligned so that a D = 16 row and its state share one line)."""
    t_off, t_key = [], []
    o = k = 0
    mult = 2 if interleaved else 1
    if interleaved:
        align = max(align, 32)
    for r, d in zip(table_rows, table_dim):
        o = (o + align - 1) // align * align
        t_off.append(o)
        t_key.append(k)
        o += r * d * mult
        k += r
    col, c = [], 0
    for t in feat_table:
        col.append(c)
        c += table_dim[t]
    return FeatureLayout(
        w_off=[t_off[t] for t in feat_table], rows=[table_rows[t] for t in feat_table],
        dim=[table_dim[t] for t in feat_table], col=col, pool=list(feat_pool),
        key_base=[t_key[t] for t in feat_table], total_keys=max(k, 1), total_dim=c,
        arena_elems=max(o, 128),   # never smaller than one (widest) row: padding slots read row 0
        stride=[table_dim[t] * 2 for t in feat_table] if interleaved else None, interleaved=interleaved)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise TzkError(f"{name}: expected a CUDA tensor, got {t.device} (no CPU fallback in torcheasyrec_b200)")
    if t.dtype != dtype:
        raise TzkError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise TzkError(f"{name}: expected a contiguous tensor")
    return t


def _rows2d(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """Accepts a 2-D fp32 CUDA tensor whose rows are contiguous (column slices of a wider buffer are fine)."""
    if not t.is_cuda:
        raise TzkError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32 or t.dim() != 2:
        raise TzkError(f"{name}: expected a 2-D float32 tensor")
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise TzkError(f"{name}: rows must be contiguous")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    return t, ld


def _unvalidated_switch(name: str) -> bool:
    """Mirrors unvalidated_switch() of csrc/tzk_common.cuh: `name`=0/1 decides, else TZK_EXPERIMENTAL=1 turns it on."""
    e = os.environ.get(name, "")
    if e[:1] in ("0", "1"):
        return e[:1] == "1"
    return os.environ.get("TZK_EXPERIMENTAL", "")[:1] == "1"


def _small_linear_rows_path(K: int, N: int) -> bool:
    """Mirrors use_bwd2() of csrc/tzk_tower.cu (launch accounting only)."""
    if os.environ.get("TZK_SMALL_LINEAR_BWD", "")[:1] == "1" or not (1 <= K <= 64 and 1 <= N <= 64):
        return False
    if os.environ.get("TZK_SMALL_LINEAR_DW", "1")[:1] != "0":
        return True
    nb = 4 if N % 4 == 0 else 1
    t = -(-N // nb) * -(-K // 4)
    if t > 128:
        t = -(-N // nb) * -(-K // 8)
    return t <= 128


def _tile_path(lay: "FeatureLayout") -> bool:
    import os

    return os.environ.get("TZK_BWD_TILE", "0") == "1" and bool(lay.vec_ok) and lay.max_dim <= 128 and not lay.interleaved


def _table_dtype(weights: torch.Tensor, name: str = "weights") -> bool:
    """Validates a table arena (fp32, or fp16 for DataType.FP16 tables); returns True for halfs."""
    if not weights.is_cuda:
        raise TzkError(f"{name}: expected a CUDA tensor, got {weights.device} (no CPU fallback in torcheasyrec_b200)")
    if weights.dtype not in (torch.float32, torch.float16):
        raise TzkError(f"{name}: expected float32 or float16 tables, got {weights.dtype}")
    if not weights.is_contiguous():
        raise TzkError(f"{name}: expected a contiguous tensor")
    return weights.dtype == torch.float16


def _opt_args(optimizer: int, state, lr: float, eps: float, ex: dict):
    """tzk_opt_args (include/tzk.h) for the _ex entry points; keeps the tensors it points to alive via the caller."""
    from ._lib import TzkOptArgs

    st2, step = ex.get("state2"), ex.get("step")
    if optimizer in (OPT_ADAM, OPT_PARTIAL_ROWWISE_ADAM) and (st2 is None or step is None):
        raise TzkError("Adam variants need state2 and the device step counter")
    for t, nm in ((st2, "state2"), (step, "step")):
        if t is not None:
            _need(t, torch.float32, nm)
    return TzkOptArgs(optimizer, lr, eps, float(ex.get("beta1", 0.9)), float(ex.get("beta2", 0.999)),
                      float(ex.get("weight_decay", 0.0)), float(ex.get("max_gradient", 0.0)),
                      _ptr(state), _ptr(st2), _ptr(step), int(bool(ex.get("weights_f16", False))),
                      int(bool(ex.get("interleaved", False))))


class CudaKernels:
    """sm_100a implementation of the hot path.  Stateless apart from cached workspaces."""

    name = "cuda"

    def __init__(self) -> None:
        self._lib = lib()
        self._ws = {}
        self.launches = 0  # hand-written tzk kernels enqueued so far (CUB's sort kernels are not counted)

    # ------------------------------------------------------------------ workspace cache
    def _workspace(self, key, nbytes: int, device) -> torch.Tensor:
        # (per host thread: two threads driving two streams must not share scratch memory)
        slot = (key, device, threading.get_ident())
        ws = self._ws.get(slot)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self._ws[slot] = ws
        return ws

    # ------------------------------------------------------------------ K3
    def lengths_to_offsets(self, lengths: torch.Tensor) -> torch.Tensor:
        _need(lengths, torch.int32, "lengths")
        n = lengths.numel()
        out = torch.empty(n + 1, dtype=torch.int64, device=lengths.device)
        nb = self._lib.tzk_lengths_to_offsets_workspace_bytes(n)
        ws = self._workspace("scan", nb, lengths.device)
        check(self._lib.tzk_lengths_to_offsets(_ptr(lengths), n, _ptr(out), _ptr(ws), ws.numel(), _stream()),
              "tzk_lengths_to_offsets")
        self.launches += 3 if n else 1
        return out

    # ------------------------------------------------------------------ K4
    def pooled_gather_fwd(self, weights: torch.Tensor, lay: FeatureLayout, ids: torch.Tensor,
                          offsets: torch.Tensor, B: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        f16 = _table_dtype(weights)
        _need(ids, torch.int64, "ids")
        _need(offsets, torch.int64, "offsets")
        F = lay.num_features
        if offsets.numel() != F * B + 1:
            raise TzkError(f"offsets has {offsets.numel()} entries, expected F*B+1 = {F * B + 1}")
        if out is None:
            out = torch.empty((B, lay.total_dim), dtype=torch.float32, device=weights.device)
        out, ld = _rows2d(out, "out")
        if lay.stride is not None:
            if f16:
                raise TzkError("strided (interleaved) tables are fp32")
            check(self._lib.tzk_pooled_gather_fwd_strided(
                _ptr(weights), _ptr(lay.d_w_off), _ptr(lay.d_rows), _ptr(lay.d_dim), _ptr(lay.d_stride),
                _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(ids), _ptr(offsets), F, B, lay.max_dim, lay.vec_ok,
                _ptr(out), ld, _stream()), "tzk_pooled_gather_fwd_strided")
            self.launches += 1
            return out
        fn = self._lib.tzk_pooled_gather_fwd_f16 if f16 else self._lib.tzk_pooled_gather_fwd
        check(fn(
            _ptr(weights), _ptr(lay.d_w_off), _ptr(lay.d_rows), _ptr(lay.d_dim), _ptr(lay.d_col),
            _ptr(lay.d_pool), _ptr(ids), _ptr(offsets), F, B, lay.max_dim, lay.vec_ok, _ptr(out), ld,
            _stream()), "tzk_pooled_gather_fwd")
        self.launches += 1
        return out

    def seq_gather_fwd(self, weights: torch.Tensor, lay: FeatureLayout, ids: torch.Tensor,
                       offsets: torch.Tensor, B: int) -> torch.Tensor:
        f16 = _table_dtype(weights)
        _need(ids, torch.int64, "ids")
        _need(offsets, torch.int64, "offsets")
        F = lay.num_features
        D = lay.dim[0] if F else 1
        if any(d != D for d in lay.dim):
            raise TzkError("seq_gather_fwd: all features of an un-pooled collection must share one dim")
        nnz = ids.numel()
        out = torch.empty((nnz, D), dtype=torch.float32, device=weights.device)
        if lay.stride is not None:
            if f16:
                raise TzkError("strided (interleaved) tables are fp32")
            check(self._lib.tzk_seq_gather_fwd_strided(_ptr(weights), _ptr(lay.d_w_off), _ptr(lay.d_rows), _ptr(ids),
                                                       _ptr(offsets), F, B, D, lay.stride[0] if F else D, nnz, _ptr(out),
                                                       _stream()), "tzk_seq_gather_fwd_strided")
            self.launches += 1
            return out
        fn = self._lib.tzk_seq_gather_fwd_f16 if f16 else self._lib.tzk_seq_gather_fwd
        check(fn(_ptr(weights), _ptr(lay.d_w_off), _ptr(lay.d_rows), _ptr(ids), _ptr(offsets), F, B, D, nnz, _ptr(out),
                 _stream()), "tzk_seq_gather_fwd")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ K5
    def fused_bwd(self, optimizer: int, pooled: bool, grad_out: torch.Tensor, weights: torch.Tensor,
                  state: Optional[torch.Tensor], lay: FeatureLayout, ids: torch.Tensor, offsets: torch.Tensor,
                  B: int, lr: float, eps: float, grad_scale: float = 1.0, **ex) -> None:
        """`ex` (optional): state2, step, beta1, beta2, weight_decay, max_gradient -> tzk_fused_bwd_ex."""
        if _table_dtype(weights):
            ex = dict(ex, weights_f16=True)
        if lay.interleaved:
            ex, state = dict(ex, interleaved=True), None    # the state rows live inside `weights`
        _need(ids, torch.int64, "ids")
        _need(offsets, torch.int64, "offsets")
        grad_out, ld = _rows2d(grad_out, "grad_out")
        if state is not None:
            _need(state, torch.float32, "state")
        F = lay.num_features
        nnz = ids.numel()
        nb = self._lib.tzk_fused_bwd_workspace_bytes(nnz, lay.total_keys, lay.max_dim)
        ws = self._workspace("bwd", nb, weights.device)
        if ex:
            oa = _opt_args(optimizer, state, lr, eps, ex)
            check(self._lib.tzk_fused_bwd_ex(
                ctypes.byref(oa), int(pooled), _ptr(grad_out), ld, _ptr(lay.d_w_off), _ptr(lay.d_rows),
                _ptr(lay.d_dim), _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(lay.d_key_base), _ptr(ids), _ptr(offsets),
                F, B, nnz, lay.total_keys, lay.max_dim, lay.vec_ok, _ptr(weights), grad_scale, _ptr(ws), ws.numel(),
                _stream()), "tzk_fused_bwd_ex")
        else:
            check(self._lib.tzk_fused_bwd(
                optimizer, int(pooled), _ptr(grad_out), ld, _ptr(lay.d_w_off), _ptr(lay.d_rows), _ptr(lay.d_dim),
                _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(lay.d_key_base), _ptr(ids), _ptr(offsets), F, B, nnz,
                lay.total_keys, lay.max_dim, lay.vec_ok, _ptr(weights), _ptr(state), lr, eps, grad_scale,
                _ptr(ws), ws.numel(), _stream()), "tzk_fused_bwd")
        # own launches next to CUB's radix sort: linearize, zero_counters, find_long_runs + the gradient half:
        # fused_apply (short runs and the long-run chunk CTAs in ONE launch), or tile_update + carry_combine
        self.launches += 5 if _tile_path(lay) else 4

    def fused_bwd_workspace_bytes(self, lay: FeatureLayout, nnz: int) -> int:
        return int(self._lib.tzk_fused_bwd_workspace_bytes(nnz, lay.total_keys, lay.max_dim))

    def fused_bwd_sort(self, pooled: bool, lay: FeatureLayout, ids: torch.Tensor, offsets: torch.Tensor, B: int,
                       ws: torch.Tensor) -> None:
        """First half of fused_bwd (linearize + radix sort of (table,row) keys): needs only the ids, so callers run
        it on a side stream while the forward pass is still going.  `ws` must stay untouched until fused_bwd_apply."""
        _need(ids, torch.int64, "ids")
        _need(offsets, torch.int64, "offsets")
        nnz = ids.numel()
        if ws.numel() < self.fused_bwd_workspace_bytes(lay, nnz):
            raise TzkError("fused_bwd_sort: workspace too small")
        check(self._lib.tzk_fused_bwd_sort(int(pooled), _ptr(lay.d_rows), _ptr(lay.d_key_base), _ptr(ids),
                                           _ptr(offsets), lay.num_features, B, nnz, lay.total_keys, lay.max_dim,
                                           _ptr(ws), ws.numel(), _stream()), "tzk_fused_bwd_sort")
        self.launches += 3

    def fused_bwd_apply(self, optimizer: int, pooled: bool, grad_out: torch.Tensor, weights: torch.Tensor,
                        state: Optional[torch.Tensor], lay: FeatureLayout, offsets: torch.Tensor, nnz: int, B: int,
                        lr: float, eps: float, grad_scale: float, ws: torch.Tensor, **ex) -> None:
        if _table_dtype(weights):
            ex = dict(ex, weights_f16=True)
        if lay.interleaved:
            ex, state = dict(ex, interleaved=True), None    # the state rows live inside `weights`
        _need(offsets, torch.int64, "offsets")
        grad_out, ld = _rows2d(grad_out, "grad_out")
        if optimizer == OPT_ACCUM_OUT:
            _need(state, torch.int32, "state (row flags)")
            ex = dict(ex, max_gradient=0.0)             # -> the _ex entry point
        elif state is not None:
            _need(state, torch.float32, "state")
        if ex:
            oa = _opt_args(optimizer, state, lr, eps, ex)
            check(self._lib.tzk_fused_bwd_apply_ex(
                ctypes.byref(oa), int(pooled), _ptr(grad_out), ld, _ptr(lay.d_w_off), _ptr(lay.d_rows),
                _ptr(lay.d_dim), _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(lay.d_key_base), _ptr(offsets),
                lay.num_features, B, nnz, lay.total_keys, lay.max_dim, lay.vec_ok, _ptr(weights), grad_scale,
                _ptr(ws), ws.numel(), _stream()), "tzk_fused_bwd_apply_ex")
        else:
            check(self._lib.tzk_fused_bwd_apply(
                optimizer, int(pooled), _ptr(grad_out), ld, _ptr(lay.d_w_off), _ptr(lay.d_rows), _ptr(lay.d_dim),
                _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(lay.d_key_base), _ptr(offsets), lay.num_features, B, nnz,
                lay.total_keys, lay.max_dim, lay.vec_ok, _ptr(weights), _ptr(state), lr, eps, grad_scale,
                _ptr(ws), ws.numel(), _stream()), "tzk_fused_bwd_apply")
        self.launches += 2 if _tile_path(lay) else 1

    # ------------------------------------------------------------------ K1 / K2
    def bucketize_rw(self, ids: torch.Tensor, offsets: torch.Tensor, F: int, B: int, W: int,
                     feat_block: torch.Tensor, want_pos: bool = False, feat_owner: Optional[torch.Tensor] = None,
                     want_inv: bool = False, wire_capacity: int = 0):
        """-> (out_lengths [W*F*B], out_offsets [W*F*B+1], out_ids [nnz], out_pos|None, out_inv|None).
        wire_capacity = C > 0: out_ids / out_pos have W*C slots (zero-filled), destination r starts at r*C."""
        _need(ids, torch.int64, "ids")
        _need(offsets, torch.int64, "offsets")
        _need(feat_block, torch.int64, "feat_block")
        if feat_owner is not None:
            _need(feat_owner, torch.int32, "feat_owner")
        dev = offsets.device
        nnz = ids.numel()
        out_lengths = torch.empty(W * F * B, dtype=torch.int32, device=dev)
        out_offsets = torch.empty(W * F * B + 1, dtype=torch.int64, device=dev)
        n_out = W * wire_capacity if wire_capacity else nnz
        out_ids = (torch.zeros if wire_capacity else torch.empty)(n_out, dtype=torch.int64, device=dev)
        out_pos = (torch.zeros if wire_capacity else torch.empty)(n_out, dtype=torch.int32, device=dev) \
            if want_pos else None
        out_inv = torch.empty(nnz, dtype=torch.int32, device=dev) if want_inv else None
        nb = self._lib.tzk_bucketize_rw_workspace_bytes(F, B, W, nnz)
        ws = self._workspace("bucketize", nb, dev)
        check(self._lib.tzk_bucketize_rw(_ptr(ids), _ptr(offsets), F, B, W, _ptr(feat_block), _ptr(feat_owner), nnz,
                                         wire_capacity, _ptr(out_lengths), _ptr(out_offsets), _ptr(out_ids), _ptr(out_pos),
                                         _ptr(out_inv), _ptr(ws), ws.numel(), _stream()), "tzk_bucketize_rw")
        self.launches += 5
        return out_lengths, out_offsets, out_ids, out_pos, out_inv

    def bag_grad_expand(self, grad_out: torch.Tensor, lay: FeatureLayout, offsets: torch.Tensor, slot: torch.Tensor,
                        B: int, n_rows: int, zero: bool = False) -> torch.Tensor:
        """g_rows[slot[l]] = grad_out[b, col_f:+D] (/L for MEAN) for every id position l of bag (f,b)."""
        grad_out, ld = _rows2d(grad_out, "grad_out")
        _need(offsets, torch.int64, "offsets")
        _need(slot, torch.int32, "slot")
        D = lay.dim[0]
        if any(d != D for d in lay.dim):
            raise TzkError("bag_grad_expand: all features must share one dim")
        out = (torch.zeros if zero else torch.empty)((n_rows, D), dtype=torch.float32, device=grad_out.device)
        check(self._lib.tzk_bag_grad_expand(_ptr(grad_out), ld, _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(offsets),
                                            _ptr(slot), lay.num_features, B, D, _ptr(out), _stream()),
              "tzk_bag_grad_expand")
        self.launches += 1
        return out

    def permute_lengths(self, lengths: torch.Tensor, perm: torch.Tensor, B: int) -> torch.Tensor:
        _need(lengths, torch.int32, "lengths")
        _need(perm, torch.int32, "perm")
        S = perm.numel()
        out = torch.empty(S * B, dtype=torch.int32, device=lengths.device)
        check(self._lib.tzk_permute_lengths(_ptr(lengths), _ptr(perm), S, B, _ptr(out), _stream()),
              "tzk_permute_lengths")
        self.launches += 1
        return out

    def permute_ids(self, ids: torch.Tensor, in_offsets: torch.Tensor, out_offsets: torch.Tensor,
                    perm: torch.Tensor, B: int, out_nnz: int) -> torch.Tensor:
        _need(ids, torch.int64, "ids")
        _need(in_offsets, torch.int64, "in_offsets")
        _need(out_offsets, torch.int64, "out_offsets")
        _need(perm, torch.int32, "perm")
        out = torch.empty(out_nnz, dtype=torch.int64, device=ids.device)
        check(self._lib.tzk_permute_ids(_ptr(ids), _ptr(in_offsets), _ptr(out_offsets), _ptr(perm),
                                        perm.numel(), B, _ptr(out), _stream()), "tzk_permute_ids")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ DIN attention over jagged rows (tzk_din.cu)
    def din_attn_input_fwd(self, query: torch.Tensor, seq: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        query, ld_q = _rows2d(query, "query")
        _need(seq, torch.float32, "seq")
        _need(offsets, torch.int64, "offsets")
        B, Dq = query.shape
        N, Ds = seq.shape
        out = torch.empty((N, 4 * Ds), dtype=torch.float32, device=seq.device)
        check(self._lib.tzk_din_attn_input_fwd(_ptr(query), ld_q, Dq, _ptr(seq), _ptr(offsets), B, Ds, N, _ptr(out),
                                               _stream()), "tzk_din_attn_input_fwd")
        self.launches += 1 if N else 0
        return out

    def din_attn_input_bwd(self, d_in: torch.Tensor, query: torch.Tensor, seq: torch.Tensor, offsets: torch.Tensor):
        query, ld_q = _rows2d(query, "query")
        _need(d_in, torch.float32, "d_in")
        _need(seq, torch.float32, "seq")
        B, Dq = query.shape
        N, Ds = seq.shape
        d_query = torch.empty((B, Dq), dtype=torch.float32, device=seq.device)
        d_seq = torch.empty((N, Ds), dtype=torch.float32, device=seq.device)
        check(self._lib.tzk_din_attn_input_bwd(_ptr(d_in), _ptr(query), ld_q, Dq, _ptr(seq), _ptr(offsets), B, Ds, N,
                                               _ptr(d_query), _ptr(d_seq), _stream()), "tzk_din_attn_input_bwd")
        self.launches += 1
        return d_query, d_seq

    def jagged_softmax_wsum_fwd(self, scores: torch.Tensor, seq: torch.Tensor, offsets: torch.Tensor, max_len: int = 0):
        _need(scores, torch.float32, "scores")
        _need(seq, torch.float32, "seq")
        _need(offsets, torch.int64, "offsets")
        N, Ds = seq.shape
        B = offsets.numel() - 1
        probs = torch.empty(N, dtype=torch.float32, device=seq.device)
        out = torch.empty((B, Ds), dtype=torch.float32, device=seq.device)
        check(self._lib.tzk_jagged_softmax_wsum_fwd(_ptr(scores), _ptr(seq), _ptr(offsets), B, Ds, int(max_len), N,
                                                    _ptr(probs), _ptr(out), _stream()), "tzk_jagged_softmax_wsum_fwd")
        self.launches += 1
        return probs, out

    def jagged_softmax_wsum_bwd(self, d_out: torch.Tensor, probs: torch.Tensor, seq: torch.Tensor, offsets: torch.Tensor,
                                max_len: int = 0):
        _need(d_out, torch.float32, "d_out")
        _need(probs, torch.float32, "probs")
        N, Ds = seq.shape
        B = offsets.numel() - 1
        d_scores = torch.empty(N, dtype=torch.float32, device=seq.device)
        d_seq = torch.empty((N, Ds), dtype=torch.float32, device=seq.device)
        check(self._lib.tzk_jagged_softmax_wsum_bwd(_ptr(d_out), _ptr(probs), _ptr(seq), _ptr(offsets), B, Ds,
                                                    int(max_len), N, _ptr(d_scores), _ptr(d_seq), _stream()),
              "tzk_jagged_softmax_wsum_bwd")
        self.launches += 1 if N else 0
        return d_scores, d_seq

    # ------------------------------------------------------------------ sharded step over peer memory (tzk_peer.cu)
    # `symm` arguments: objects with `.ptrs` = ctypes array [W] of device addresses (rank r's symmetric buffer as
    # mapped in this process) — peer_exchange._Symm.
    def peer_pooled_gather_fwd(self, tables, rf_w_off: torch.Tensor, feat_rows: torch.Tensor, feat_block: torch.Tensor,
                               feat_owner: torch.Tensor, lay: FeatureLayout, ids: torch.Tensor, offsets: torch.Tensor,
                               B: int, W: int, out: Optional[torch.Tensor] = None, mirror: Optional[torch.Tensor] = None,
                               feat_mirror_off: Optional[torch.Tensor] = None,
                               feat_sel: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`feat_sel` (device int32 indices): serve only these features (their output columns); `out` is then required."""
        _need(ids, torch.int64, "ids")
        _need(offsets, torch.int64, "offsets")
        F = lay.num_features
        if offsets.numel() != F * B + 1:
            raise TzkError(f"offsets has {offsets.numel()} entries, expected F*B+1 = {F * B + 1}")
        if not lay.vec_ok:
            raise TzkError("peer gather: rows must be 16-B aligned (dims and offsets multiples of 4 floats)")
        if out is None:
            out = torch.empty((B, lay.total_dim), dtype=torch.float32, device=ids.device)
        out, ld = _rows2d(out, "out")
        if feat_sel is not None:
            _need(feat_sel, torch.int32, "feat_sel")
            check(self._lib.tzk_peer_pooled_gather_fwd_sel(
                tables.ptrs, _ptr(rf_w_off), _ptr(feat_rows), _ptr(feat_block), _ptr(feat_owner), _ptr(lay.d_dim),
                _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(ids), _ptr(offsets), F, B, W, (lay.max_dim + 3) // 4 * 4,
                _ptr(out), ld, _ptr(mirror), _ptr(feat_mirror_off), _ptr(feat_sel), feat_sel.numel(), _stream()),
                "tzk_peer_pooled_gather_fwd_sel")
            self.launches += 1
            return out
        check(self._lib.tzk_peer_pooled_gather_fwd(
            tables.ptrs, _ptr(rf_w_off), _ptr(feat_rows), _ptr(feat_block), _ptr(feat_owner), _ptr(lay.d_dim),
            _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(ids), _ptr(offsets), F, B, W, (lay.max_dim + 3) // 4 * 4, _ptr(out),
            ld, _ptr(mirror), _ptr(feat_mirror_off), _stream()), "tzk_peer_pooled_gather_fwd")
        self.launches += 1
        return out

    def peer_seq_gather_fwd(self, tables, rf_w_off: torch.Tensor, feat_rows: torch.Tensor, feat_block: torch.Tensor,
                            feat_owner: torch.Tensor, lay: FeatureLayout, ids: torch.Tensor, offsets: torch.Tensor,
                            B: int, W: int, mirror: Optional[torch.Tensor] = None,
                            feat_mirror_off: Optional[torch.Tensor] = None) -> torch.Tensor:
        _need(ids, torch.int64, "ids")
        _need(offsets, torch.int64, "offsets")
        F, D, nnz = lay.num_features, lay.dim[0], ids.numel()
        out = torch.empty((nnz, D), dtype=torch.float32, device=ids.device)
        check(self._lib.tzk_peer_seq_gather_fwd(tables.ptrs, _ptr(rf_w_off), _ptr(feat_rows), _ptr(feat_block),
                                                _ptr(feat_owner), _ptr(ids), _ptr(offsets), F, B, W, D, nnz, _ptr(out),
                                                _ptr(mirror), _ptr(feat_mirror_off), _stream()),
              "tzk_peer_seq_gather_fwd")
        self.launches += 1 if nnz else 0
        return out

    def peer_mirror_refresh(self, tables, W: int, seg_rank: torch.Tensor, seg_src: torch.Tensor, seg_dst: torch.Tensor,
                            seg_n: torch.Tensor, mirror: torch.Tensor) -> None:
        """This step's local copy of the small tables (see tzk_peer_mirror_refresh)."""
        _need(mirror, torch.float32, "mirror")
        check(self._lib.tzk_peer_mirror_refresh(tables.ptrs, W, _ptr(seg_rank), _ptr(seg_src), _ptr(seg_dst), _ptr(seg_n),
                                                seg_rank.numel(), _ptr(mirror), _stream()), "tzk_peer_mirror_refresh")
        self.launches += 1

    def peer_barrier(self, pads, me: int, W: int, epoch: torch.Tensor) -> None:
        check(self._lib.tzk_peer_barrier(pads.ptrs, me, W, _ptr(epoch), _stream()), "tzk_peer_barrier")
        self.launches += 1

    def peer_bucketize(self, ids: torch.Tensor, offsets: torch.Tensor, F: int, B: int, W: int, feat_block: torch.Tensor,
                       feat_owner: torch.Tensor, feat_rows: torch.Tensor, rf_key_base: torch.Tensor, pooled: bool,
                       cap: int, wire_key: torch.Tensor, wire_idx: torch.Tensor, counts: torch.Tensor) -> None:
        """ids of the local batch -> this rank's own wire buffers (see tzk_peer_bucketize in include/tzk.h)."""
        _need(ids, torch.int64, "ids")
        _need(offsets, torch.int64, "offsets")
        _need(wire_key, torch.int64, "wire_key")
        _need(wire_idx, torch.int32, "wire_idx")
        _need(counts, torch.int32, "counts")
        if wire_key.numel() < W * cap or wire_idx.numel() < W * cap or counts.numel() < W + 1:
            raise TzkError("peer_bucketize: wire buffers smaller than W * cap")
        nb = self._lib.tzk_peer_bucketize_workspace_bytes(F, B, W)
        ws = self._workspace(("peer_bkt", F, B, W), nb, ids.device)
        check(self._lib.tzk_peer_bucketize(_ptr(ids), _ptr(offsets), F, B, W, _ptr(feat_block), _ptr(feat_owner),
                                           _ptr(feat_rows), _ptr(rf_key_base), int(pooled), cap, _ptr(wire_key),
                                           _ptr(wire_idx), _ptr(counts), _ptr(ws), ws.numel(), _stream()),
              "tzk_peer_bucketize")
        self.launches += 3

    def peer_publish_grad(self, grad: torch.Tensor, lay: FeatureLayout, offsets: torch.Tensor, B: int,
                          dst: torch.Tensor) -> None:
        grad, ld = _rows2d(grad, "grad")
        dst, ld_dst = _rows2d(dst, "dst")
        if POOL_MEAN not in lay.pool:        # plain copy: no per-bag scale to fold in
            dst.copy_(grad)
            return
        check(self._lib.tzk_peer_publish_grad(_ptr(grad), ld, _ptr(lay.d_col), _ptr(lay.d_dim), _ptr(lay.d_pool),
                                              _ptr(offsets), lay.num_features, B, _ptr(dst), ld_dst, _stream()),
              "tzk_peer_publish_grad")
        self.launches += 1

    def peer_push_grad(self, recv, grad: torch.Tensor, lay: FeatureLayout, offsets: torch.Tensor, wire_idx: torch.Tensor,
                       counts: torch.Tensor, me: int, W: int, cap: int, B: int, pooled: bool) -> None:
        """This rank's gradient slices -> the owners' receive buffers, wire order (see tzk_peer_push_grad)."""
        grad, ld = _rows2d(grad, "grad")
        _need(wire_idx, torch.int32, "wire_idx")
        _need(counts, torch.int32, "counts")
        check(self._lib.tzk_peer_push_grad(recv.ptrs, _ptr(grad), ld, _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(offsets),
                                           _ptr(wire_idx), _ptr(counts), me, W, cap, B, lay.dim[0], int(pooled),
                                           _stream()), "tzk_peer_push_grad")
        self.launches += 1

    def peer_allreduce_mean(self, srcs, W: int, n: int, out: torch.Tensor) -> None:
        _need(out, torch.float32, "out")
        check(self._lib.tzk_peer_allreduce_mean(srcs.ptrs, W, n, _ptr(out), _stream()), "tzk_peer_allreduce_mean")
        self.launches += 1

    def peer_small_update(self, optimizer: int, psum, flags, W: int, tabs: torch.Tensor, n_tabs: int, total_rows: int,
                          max_dim: int, weights: torch.Tensor, state: Optional[torch.Tensor], lr: float, eps: float,
                          **ex) -> None:
        """Owner side of the small-table exchange (see tzk_peer_small_update): psum / flags are symmetric buffers."""
        _need(weights, torch.float32, "weights")
        oa = _opt_args(optimizer, state, lr, eps, ex)
        check(self._lib.tzk_peer_small_update(ctypes.byref(oa), psum.ptrs, flags.ptrs, W, _ptr(tabs), n_tabs, total_rows,
                                              max_dim, _ptr(weights), _stream()), "tzk_peer_small_update")
        self.launches += 1

    def fused_bwd_sort_peer(self, wire_key, wire_idx, counts, me: int, W: int, cap: int, idx_span: int,
                            lay: FeatureLayout, overflow: Optional[torch.Tensor], ws: torch.Tensor) -> None:
        if ws.numel() < self.fused_bwd_workspace_bytes(lay, W * cap):
            raise TzkError("fused_bwd_sort_peer: workspace too small")
        # idx_span == 0: "slot mode" (the gradient rows are pushed to this rank's receive buffer: value = its row)
        check(self._lib.tzk_fused_bwd_sort_peer(wire_key.ptrs, wire_idx.ptrs if idx_span else None, counts.ptrs, me, W,
                                                cap, idx_span,
                                                lay.total_keys, lay.max_dim, _ptr(overflow), _ptr(ws), ws.numel(),
                                                _stream()), "tzk_fused_bwd_sort_peer")
        self.launches += 3

    def fused_bwd_apply_peer(self, optimizer: int, pooled: bool, grads, ld_grad: int, weights: torch.Tensor,
                             state: Optional[torch.Tensor], lay: FeatureLayout, B: int, me: int, W: int, cap: int,
                             idx_span: int, lr: float, eps: float, grad_scale: float, ws: torch.Tensor, **ex) -> None:
        _need(weights, torch.float32, "weights")
        if state is not None:
            _need(state, torch.float32, "state")
        oa = _opt_args(optimizer, state, lr, eps, ex)
        check(self._lib.tzk_fused_bwd_apply_peer(
            ctypes.byref(oa), int(pooled), grads.ptrs, ld_grad, _ptr(lay.d_w_off), _ptr(lay.d_rows), _ptr(lay.d_dim),
            _ptr(lay.d_col), _ptr(lay.d_pool), _ptr(lay.d_key_base), lay.num_features, B, me, W, cap, idx_span,
            lay.total_keys, lay.max_dim, lay.vec_ok, _ptr(weights), grad_scale, _ptr(ws), ws.numel(), _stream()),
            "tzk_fused_bwd_apply_peer")
        self.launches += 2 if _tile_path(lay) else 1

    # ------------------------------------------------------------------ K6
    def col_gather_sum(self, srcs: Sequence[torch.Tensor], plan: "ColPlan", rows: int,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
        lds = []
        for i, s in enumerate(srcs):
            s, ld = _rows2d(s, f"srcs[{i}]")
            lds.append(ld)
        dev = srcs[0].device
        if out is None:
            out = torch.empty((rows, plan.C), dtype=torch.float32, device=dev)
        out, ld_out = _rows2d(out, "out")
        n = len(srcs)
        ptrs = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])  # host arrays -> kernel parameters
        ldt = (ctypes.c_int64 * n)(*lds)
        check(self._lib.tzk_col_gather_sum(ptrs, ldt, n, _ptr(plan.d_col_start), _ptr(plan.d_col_src),
                                           _ptr(plan.d_col_srccol), plan.C, rows, _ptr(out), ld_out, _stream()),
              "tzk_col_gather_sum")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ K7
    def jagged_to_padded(self, values: torch.Tensor, offsets: torch.Tensor, T: int) -> torch.Tensor:
        _need(values, torch.float32, "values")
        _need(offsets, torch.int64, "offsets")
        B = offsets.numel() - 1
        D = values.shape[1]
        out = torch.empty((B, T, D), dtype=torch.float32, device=values.device)
        check(self._lib.tzk_jagged_to_padded(_ptr(values), _ptr(offsets), B, T, D, _ptr(out), _stream()),
              "tzk_jagged_to_padded")
        self.launches += 1
        return out

    def padded_to_jagged(self, grad_out: torch.Tensor, offsets: torch.Tensor, nnz: int) -> torch.Tensor:
        _need(grad_out, torch.float32, "grad_out")
        _need(offsets, torch.int64, "offsets")
        B, T, D = grad_out.shape
        out = torch.empty((nnz, D), dtype=torch.float32, device=grad_out.device)
        check(self._lib.tzk_padded_to_jagged(_ptr(grad_out), _ptr(offsets), B, T, D, nnz, _ptr(out), _stream()),
              "tzk_padded_to_jagged")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ A7
    def fm_fwd(self, x: torch.Tensor, N: int, D: int) -> torch.Tensor:
        x, ld = _rows2d(x, "x")
        B = x.shape[0]
        y = torch.empty((B, D), dtype=torch.float32, device=x.device)
        check(self._lib.tzk_fm_fwd(_ptr(x), ld, B, N, D, _ptr(y), D, _stream()), "tzk_fm_fwd")
        self.launches += 1
        return y

    def fm_bwd(self, x: torch.Tensor, dy: torch.Tensor, N: int, D: int) -> torch.Tensor:
        x, ld = _rows2d(x, "x")
        dy, ld_dy = _rows2d(dy, "dy")
        B = x.shape[0]
        dx = torch.empty((B, N * D), dtype=torch.float32, device=x.device)
        check(self._lib.tzk_fm_bwd(_ptr(x), ld, _ptr(dy), ld_dy, B, N, D, _ptr(dx), N * D, _stream()),
              "tzk_fm_bwd")
        self.launches += 1
        return dx

    # ------------------------------------------------------------------ A9 / A10
    def dot_interact_fwd(self, dense: Optional[torch.Tensor], sparse: torch.Tensor, Ns: int, D: int,
                         copy_dense: bool, copy_sparse: bool, pad_to: int = 1, p_pad: int = 0) -> torch.Tensor:
        """Layout [P | p_pad zeros | dense D | sparse Ns*D | tail zeros up to a multiple of pad_to]."""
        sparse, ld_s = _rows2d(sparse, "sparse")
        B = sparse.shape[0]
        ld_d = 0
        if dense is not None:
            dense, ld_d = _rows2d(dense, "dense")
        N = Ns + (dense is not None)
        width = N * (N - 1) // 2 + p_pad + (D if (copy_dense and dense is not None) else 0) + \
            (Ns * D if copy_sparse else 0)
        wp = (width + pad_to - 1) // pad_to * pad_to
        out = torch.empty((B, wp), dtype=torch.float32, device=sparse.device)
        if wp != width:
            out[:, width:].zero_()
        check(self._lib.tzk_dot_interact_fwd(_ptr(dense), ld_d, _ptr(sparse), ld_s, B, Ns, D, int(copy_dense),
                                             int(copy_sparse), p_pad, _ptr(out), wp, _stream()),
              "tzk_dot_interact_fwd")
        self.launches += 1
        return out

    def dot_interact_bwd(self, dense: Optional[torch.Tensor], sparse: torch.Tensor, d_out: torch.Tensor,
                         Ns: int, D: int, copy_dense: bool, copy_sparse: bool, p_pad: int = 0):
        sparse, ld_s = _rows2d(sparse, "sparse")
        d_out, ld_o = _rows2d(d_out, "d_out")
        B = sparse.shape[0]
        ld_d = 0
        d_dense = None
        if dense is not None:
            dense, ld_d = _rows2d(dense, "dense")
            d_dense = torch.empty((B, D), dtype=torch.float32, device=sparse.device)
        d_sparse = torch.empty((B, Ns * D), dtype=torch.float32, device=sparse.device)
        check(self._lib.tzk_dot_interact_bwd(_ptr(dense), ld_d, _ptr(sparse), ld_s, _ptr(d_out), ld_o, B, Ns, D,
                                             int(copy_dense), int(copy_sparse), p_pad, _ptr(d_dense), D,
                                             _ptr(d_sparse), Ns * D, _stream()), "tzk_dot_interact_bwd")
        self.launches += 1
        return d_dense, d_sparse

    # ------------------------------------------------------------------ dense-tower helpers
    def bias_act(self, y: torch.Tensor, bias: Optional[torch.Tensor], relu: bool) -> torch.Tensor:
        y, ld = _rows2d(y, "y")
        check(self._lib.tzk_bias_act(_ptr(y), ld, _ptr(bias), y.shape[0], y.shape[1], int(relu), _stream()),
              "tzk_bias_act")
        self.launches += 1
        return y

    def act_bwd_colsum(self, dy: torch.Tensor, y: Optional[torch.Tensor], relu: bool, want_dz: bool = True):
        dy, ld_dy = _rows2d(dy, "dy")
        M, N = dy.shape
        ld_y = 0
        if y is not None:
            y, ld_y = _rows2d(y, "y")
        dz = torch.empty((M, N), dtype=torch.float32, device=dy.device) if want_dz else None
        colsum = torch.empty(N, dtype=torch.float32, device=dy.device)
        nb = self._lib.tzk_act_bwd_colsum_workspace_bytes(M, N)
        ws = self._workspace("colsum", nb, dy.device)
        check(self._lib.tzk_act_bwd_colsum(_ptr(dy), ld_dy, _ptr(y), ld_y, M, N, int(relu), _ptr(dz), N, _ptr(colsum),
                                           _ptr(ws), ws.numel(), _stream()), "tzk_act_bwd_colsum")
        self.launches += 2
        return dz, colsum
    # ------------------------------------------------------------------ narrow layers + BCE head
    def small_linear_fwd(self, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor],
                         relu: bool) -> torch.Tensor:
        x, ld_x = _rows2d(x, "x")
        _need(w, torch.float32, "w")
        M, K = x.shape
        N = w.shape[0]
        if w.shape[1] != K:
            raise TzkError(f"small_linear_fwd: x has {K} columns, w expects {w.shape[1]}")
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        if M:
            check(self._lib.tzk_small_linear_fwd(_ptr(x), ld_x, _ptr(w), _ptr(bias), M, K, N, int(relu), _ptr(y), N,
                                                 _stream()), "tzk_small_linear_fwd")
            self.launches += 1
        return y

    def small_linear_bwd(self, x: torch.Tensor, w: torch.Tensor, y: Optional[torch.Tensor], dy: torch.Tensor,
                         relu: bool, want_dx: bool, want_db: bool):
        """-> (dx | None, dw [N,K], db [N] | None)."""
        x, ld_x = _rows2d(x, "x")
        dy, ld_dy = _rows2d(dy, "dy")
        _need(w, torch.float32, "w")
        M, K = x.shape
        N = w.shape[0]
        ld_y = 0
        if relu:
            y, ld_y = _rows2d(y, "y")
        dx = torch.empty((M, K), dtype=torch.float32, device=x.device) if want_dx else None
        dw = torch.empty((N, K), dtype=torch.float32, device=x.device)
        db = torch.empty(N, dtype=torch.float32, device=x.device) if want_db else None
        nb = self._lib.tzk_small_linear_bwd_workspace_bytes(M, K, N)
        ws = self._workspace(("slb", K, N), nb, x.device)
        check(self._lib.tzk_small_linear_bwd(_ptr(x), ld_x, _ptr(w), _ptr(y) if relu else None, ld_y, _ptr(dy), ld_dy,
                                             M, K, N, int(relu), _ptr(dx), K, _ptr(dw), _ptr(db), _ptr(ws),
                                             ws.numel(), _stream()), "tzk_small_linear_bwd")
        # tile kernel + reduction, or (tzk_tower_bwd2.cuh) dx rows kernel + dW kernel + reduction
        self.launches += (2 + int(want_dx)) if _small_linear_rows_path(K, N) else 2
        return dx, dw, db

    def bce_logits_fwd_bwd(self, logits: torch.Tensor, labels: torch.Tensor, want_grad: bool = True):
        """-> (loss [scalar tensor], dloss/dlogits [M] | None); mean reduction."""
        _need(logits, torch.float32, "logits")
        _need(labels, torch.float32, "labels")
        M = logits.numel()
        if labels.numel() != M:
            raise TzkError("bce_logits: logits and labels differ in size")
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dz = torch.empty(M, dtype=torch.float32, device=logits.device) if want_grad else None
        nb = self._lib.tzk_bce_logits_workspace_bytes(M)
        ws = self._workspace("bce", nb, logits.device)
        check(self._lib.tzk_bce_logits_fwd_bwd(_ptr(logits), _ptr(labels), M, _ptr(loss), _ptr(dz), _ptr(ws),
                                               ws.numel(), _stream()), "tzk_bce_logits_fwd_bwd")
        self.launches += 2
        return loss, dz

    def tower_tail_bce(self, y1: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor,
                       b2: Optional[torch.Tensor], labels: torch.Tensor):
        """Last Perceptron (K -> N, ReLU) + Linear(N, 1) + mean BCE, forward and backward in one pass
        (csrc/tzk_tower_tail.cuh).  -> (loss [scalar], logits [M], dy1 [M, K], dW1 [N, K], db1 [N], dw2 [1, N], db2 [1])."""
        y1, ld = _rows2d(y1, "y1")
        _need(w1, torch.float32, "w1")
        _need(w2, torch.float32, "w2")
        _need(labels, torch.float32, "labels")
        M, K = y1.shape
        N = w1.shape[0]
        if w1.shape != (N, K) or w2.numel() != N or labels.numel() != M:
            raise TzkError("tower_tail_bce: shapes do not chain")
        dev = y1.device
        logits = torch.empty(M, dtype=torch.float32, device=dev)
        dy1 = torch.empty((M, K), dtype=torch.float32, device=dev)
        out = torch.empty(N * K + 2 * N + 2, dtype=torch.float32, device=dev)
        nb = self._lib.tzk_tower_tail_bce_workspace_bytes(M, K, N)
        ws = self._workspace("tail", nb, dev)
        check(self._lib.tzk_tower_tail_bce(_ptr(y1), ld, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(labels), M, K, N,
                                           _ptr(logits), _ptr(dy1), K, _ptr(out), _ptr(ws), ws.numel(), _stream()),
              "tzk_tower_tail_bce")
        self.launches += 2
        o = N * K
        return (out[o + 2 * N + 1], logits, dy1, out[:o].view(N, K), out[o:o + N], out[o + N:o + 2 * N].view(1, N),
                out[o + 2 * N:o + 2 * N + 1])


@dataclass
class ColPlan:
    """CSR description of a column gather-sum (K6): column c of the destination sums
    srcs[col_src[k]][:, col_srccol[k]] for k in [col_start[c], col_start[c+1])."""

    col_start: List[int]
    col_src: List[int]
    col_srccol: List[int]
    d_col_start: Optional[torch.Tensor] = None
    d_col_src: Optional[torch.Tensor] = None
    d_col_srccol: Optional[torch.Tensor] = None

    @property
    def C(self) -> int:
        return len(self.col_start) - 1

    def to(self, device) -> "ColPlan":
        self.d_col_start = torch.tensor(self.col_start, dtype=torch.int32, device=device)
        self.d_col_src = torch.tensor(self.col_src or [0], dtype=torch.int32, device=device)
        self.d_col_srccol = torch.tensor(self.col_srccol or [0], dtype=torch.int32, device=device)
        return self


_default: Optional[CudaKernels] = None


def default_kernels() -> CudaKernels:
    global _default
    if _default is None:
        _default = CudaKernels()
    return _default
