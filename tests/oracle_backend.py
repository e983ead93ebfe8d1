"""Checker backend: the CudaKernels method set implemented with the CPU oracle (numpy).

TEST INFRASTRUCTURE.  Lets tests drive the package's HOST logic (EmbeddingGroup, regroup plans, sharding /
all-to-all plumbing under gloo) on a box without a GPU, and gives -m gpu tests an independent model-level
reference.  Never imported by torcheasyrec_b200 itself.
"""
import numpy as np
import torch

from oracle import tzk_oracle as O


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _tables(weights, lay):
    """Views of the arena (shared memory with the torch tensor) as per-table arrays, one per feature slot."""
    arr = weights.detach().numpy() if not weights.is_cuda else None
    assert arr is not None, "oracle backend works on CPU tensors"
    tabs, feat_table, seen = [], [], {}
    for f in range(lay.num_features):
        key = (lay.w_off[f], lay.rows[f], lay.dim[f])   # zero-row shards share an offset with their neighbour
        if key not in seen:
            seen[key] = len(tabs)
            tabs.append(arr[key[0]:key[0] + lay.rows[f] * lay.dim[f]].reshape(lay.rows[f], lay.dim[f]))
        feat_table.append(seen[key])
    return tabs, feat_table


def _c(a):
    return np.ascontiguousarray(_np(a)) if a is not None else None


class OracleKernels:
    """use_c=True routes the heavy steps through oracle/libtzk_oracle.so (C/OpenMP, same arithmetic); the CPU
    baseline of bench.py uses that, parity tests default to the numpy restatement."""

    name = "oracle"

    def __init__(self, use_c: bool = False) -> None:
        self.use_c = False
        if use_c:
            from oracle import c_oracle

            if c_oracle.available():
                self.use_c, self.C = True, c_oracle
                self.name = "oracle-c"

    def lengths_to_offsets(self, lengths):
        return torch.from_numpy(O.lengths_to_offsets(_np(lengths)))

    def pooled_gather_fwd(self, weights, lay, ids, offsets, B, out=None):
        if self.use_c:
            res = torch.from_numpy(self.C.pooled_lookup(_c(weights), lay, _c(ids), _c(offsets), B))
        else:
            tabs, ft = _tables(weights, lay)
            res = torch.from_numpy(O.pooled_lookup(tabs, ft, lay.pool, _np(ids), _np(offsets), B))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def seq_gather_fwd(self, weights, lay, ids, offsets, B):
        tabs, ft = _tables(weights, lay)
        return torch.from_numpy(O.seq_lookup(tabs, ft, _np(ids), _np(offsets), B))

    def fused_bwd(self, optimizer, pooled, grad_out, weights, state, lay, ids, offsets, B, lr, eps, grad_scale=1.0,
                  **ex):
        if self.use_c and pooled and not ex:
            self.C.fused_update(optimizer, _np(grad_out), weights.detach().numpy(),
                                None if state is None else state.numpy(), lay, _c(ids), _c(offsets), B, lr, eps,
                                grad_scale)
            return
        tabs, ft = _tables(weights, lay)
        states = [None] * len(tabs)
        if state is not None:
            sarr = state.numpy()
            for f in range(lay.num_features):
                t = ft[f]
                if states[t] is None:
                    if optimizer in (O.OPT_ADAGRAD, O.OPT_ADAM, O.OPT_PARTIAL_ROWWISE_ADAM):
                        states[t] = sarr[lay.w_off[f]:lay.w_off[f] + lay.rows[f] * lay.dim[f]].reshape(lay.rows[f], lay.dim[f])
                    else:
                        states[t] = sarr[lay.key_base[f]:lay.key_base[f] + lay.rows[f]]
        states2, kw = None, {}
        if ex:
            kw = dict(beta1=ex.get("beta1", 0.9), beta2=ex.get("beta2", 0.999), weight_decay=ex.get("weight_decay", 0.0),
                      max_gradient=ex.get("max_gradient", 0.0))
            if ex.get("state2") is not None:
                s2 = ex["state2"].numpy()
                states2 = [None] * len(tabs)
                for f in range(lay.num_features):
                    t = ft[f]
                    if states2[t] is None:
                        if optimizer == O.OPT_ADAM:
                            states2[t] = s2[lay.w_off[f]:lay.w_off[f] + lay.rows[f] * lay.dim[f]].reshape(lay.rows[f], lay.dim[f])
                        else:
                            states2[t] = s2[lay.key_base[f]:lay.key_base[f] + lay.rows[f]]
                kw["step"] = int(round(float(ex["step"])))
        O.fused_update(optimizer, tabs, states, ft, lay.pool, _np(ids), _np(offsets), B, _np(grad_out), lr, eps,
                       grad_scale, pooled=bool(pooled), states2=states2, **kw)

    def bucketize_rw(self, ids, offsets, F, B, W, feat_block, want_pos=False, feat_owner=None, want_inv=False,
                     wire_capacity=0):
        ol, oo, oi, op = O.bucketize_rw(_np(ids), _np(offsets), F, B, W, _np(feat_block).tolist(),
                                        None if feat_owner is None else _np(feat_owner).tolist())
        inv = np.empty(len(op), dtype=np.int32)
        inv[op] = np.arange(len(op), dtype=np.int32)
        if wire_capacity:      # fixed-capacity wire layout: destination r starts at r*C
            C = wire_capacity
            dest_start = oo[::F * B]
            slot = np.arange(len(oi))
            r_of = np.searchsorted(dest_start[1:], slot, side="right").clip(max=W - 1)
            rel = slot - dest_start[r_of]
            keep = rel < C
            pslot = (r_of * C + rel)
            pid = np.zeros(W * C, dtype=np.int64)
            ppos = np.zeros(W * C, dtype=np.int32)
            pid[pslot[keep]] = oi[keep]
            ppos[pslot[keep]] = op[keep]
            inv = np.where(keep, pslot, r_of * C)[inv].astype(np.int32)
            oi, op = pid, ppos
        inv = torch.from_numpy(inv) if want_inv else None
        return (torch.from_numpy(ol), torch.from_numpy(oo), torch.from_numpy(oi),
                torch.from_numpy(op) if want_pos else None, inv)

    def bag_grad_expand(self, grad_out, lay, offsets, slot, B, n_rows, zero=False):
        return torch.from_numpy(O.bag_grad_expand(_np(grad_out), lay.col, lay.pool, _np(offsets), _np(slot),
                                                  lay.num_features, B, lay.dim[0], n_rows))

    def permute_lengths(self, lengths, perm, B):
        l = _np(lengths)
        p = _np(perm)
        return torch.from_numpy(np.concatenate([l[i * B:(i + 1) * B] for i in p]).astype(np.int32)
                                if len(p) else l[:0])

    def permute_ids(self, ids, in_offsets, out_offsets, perm, B, out_nnz):
        i, off, p = _np(ids), _np(in_offsets), _np(perm)
        parts = [i[off[k * B]:off[(k + 1) * B]] for k in p]
        return torch.from_numpy(np.concatenate(parts).astype(np.int64) if parts else i[:0])

    def col_gather_sum(self, srcs, plan, rows, out=None):
        res = np.zeros((rows, plan.C), dtype=np.float32)
        arrs = [_np(s) for s in srcs]
        for c in range(plan.C):
            for k in range(plan.col_start[c], plan.col_start[c + 1]):
                res[:, c] += arrs[plan.col_src[k]][:, plan.col_srccol[k]]
        res = torch.from_numpy(res)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def jagged_to_padded(self, values, offsets, T):
        return torch.from_numpy(O.to_padded_dense(_np(values), _np(offsets), T))

    def padded_to_jagged(self, grad_out, offsets, nnz):
        return torch.from_numpy(O.padded_to_jagged(_np(grad_out), _np(offsets), nnz))

    def fm_fwd(self, x, N, D):
        if self.use_c:
            return torch.from_numpy(self.C.fm_fwd(_c(x), N, D))
        return torch.from_numpy(O.fm(_np(x).reshape(-1, N, D)))

    def fm_bwd(self, x, dy, N, D):
        if self.use_c:
            return torch.from_numpy(self.C.fm_bwd(_c(x), _c(dy), N, D))
        return torch.from_numpy(O.fm_bwd(_np(x).reshape(-1, N, D), _np(dy)).reshape(-1, N * D))

    def dot_interact_fwd(self, dense, sparse, Ns, D, copy_dense, copy_sparse, pad_to=1, p_pad=0):
        if self.use_c:
            res = self.C.dot_interact_fwd(_c(dense), _c(sparse), Ns, D, copy_dense, copy_sparse)
        else:
            res = O.dlrm_interact(_np(dense), _np(sparse), Ns, D, copy_dense, copy_sparse)
        if p_pad:
            N = Ns + (dense is not None)
            P = N * (N - 1) // 2
            res = np.concatenate([res[:, :P], np.zeros((res.shape[0], p_pad), np.float32), res[:, P:]], axis=1)
        pad = (-res.shape[1]) % pad_to
        if pad:
            res = np.concatenate([res, np.zeros((res.shape[0], pad), np.float32)], axis=1)
        return torch.from_numpy(np.ascontiguousarray(res))

    def dot_interact_bwd(self, dense, sparse, d_out, Ns, D, copy_dense, copy_sparse, p_pad=0):
        if p_pad:
            N = Ns + (dense is not None)
            P = N * (N - 1) // 2
            d_out = torch.cat([d_out[:, :P], d_out[:, P + p_pad:]], dim=1)
        if self.use_c:
            dd, ds = self.C.dot_interact_bwd(_c(dense), _c(sparse), _c(d_out), Ns, D, copy_dense, copy_sparse)
            return (None if dd is None else torch.from_numpy(dd)), torch.from_numpy(ds)
        dd, ds = O.dlrm_interact_bwd(_np(dense), _np(sparse), _np(d_out), Ns, D, copy_dense, copy_sparse)
        return (None if dd is None else torch.from_numpy(dd)), torch.from_numpy(ds)
