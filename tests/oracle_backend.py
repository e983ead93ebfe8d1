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
            st = lay.row_stride(f)      # interleaved arenas: [weight row | state row] lines -> a strided view
            tabs.append(arr[key[0]:key[0] + lay.rows[f] * st].reshape(lay.rows[f], st)[:, :lay.dim[f]])
        feat_table.append(seen[key])
    return tabs, feat_table


def _interleaved_states(weights, lay, ft, n_tabs):
    """The accumulator halves of an interleaved arena, one strided view per table."""
    arr = weights.detach().numpy()
    states = [None] * n_tabs
    for f in range(lay.num_features):
        t = ft[f]
        if states[t] is None:
            st, d = lay.row_stride(f), lay.dim[f]
            states[t] = arr[lay.w_off[f]:lay.w_off[f] + lay.rows[f] * st].reshape(lay.rows[f], st)[:, d:]
    return states


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
        if self.use_c and lay.stride is None:
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
        if self.use_c and pooled and not ex and lay.stride is None:
            self.C.fused_update(optimizer, _np(grad_out), weights.detach().numpy(),
                                None if state is None else state.numpy(), lay, _c(ids), _c(offsets), B, lr, eps,
                                grad_scale)
            return
        tabs, ft = _tables(weights, lay)
        states = [None] * len(tabs)
        if lay.interleaved:
            states = _interleaved_states(weights, lay, ft, len(tabs))
        elif state is not None:
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

    # ------------------------------------------------------------------ sharded step over peer memory (model of tzk_peer.cu)
    # `symm` arguments are the tests' in-process stand-ins for symmetric allocations: `.everyone[r]` = rank r's tensor.
    @staticmethod
    def _owner_of(i, block, owner, W):
        q = i // block
        r = owner + q
        if r >= W:
            q -= r - (W - 1)
            r = W - 1
        return r, i - q * block

    def peer_mirror_refresh(self, tables, W, seg_rank, seg_src, seg_dst, seg_n, mirror):
        m = mirror.numpy()
        for r, s, d, n in zip(seg_rank.tolist(), seg_src.tolist(), seg_dst.tolist(), seg_n.tolist()):
            m[d:d + n] = tables.everyone[r].numpy()[s:s + n]

    def peer_pooled_gather_fwd(self, tables, rf_w_off, feat_rows, feat_block, feat_owner, lay, ids, offsets, B, W,
                               out=None, mirror=None, feat_mirror_off=None, feat_sel=None):
        sel = None if feat_sel is None else set(feat_sel.tolist())
        uses_mirror = mirror is not None and (sel is None or any(feat_mirror_off.tolist()[f] >= 0 for f in sel))
        if uses_mirror:              # the model keeps reading the owners (the mirror is checked to be an exact copy)
            self._check_mirror(tables, rf_w_off, feat_rows, feat_block, feat_owner, lay, W, mirror, feat_mirror_off)
        F = lay.num_features
        blocks, owners = feat_block.tolist(), feat_owner.tolist()
        rows, w_off = feat_rows.tolist(), rf_w_off.tolist()
        idl, off = ids.tolist(), offsets.tolist()
        o = np.zeros((B, lay.total_dim), dtype=np.float32)
        if sel is not None:          # a launch over a feature list writes those features' columns only
            assert out is not None
            o = out.numpy()
        for f in range(F):
            if sel is not None and f not in sel:
                continue
            D, col = lay.dim[f], lay.col[f]
            for b in range(B):
                s, e = off[f * B + b], off[f * B + b + 1]
                acc = np.zeros(D, dtype=np.float32)
                for l in range(s, e):
                    i = idl[l] if 0 <= idl[l] < rows[f] else 0
                    r, loc = self._owner_of(i, blocks[f], owners[f], W)
                    base = w_off[r * F + f] + loc * D
                    acc = acc + tables.everyone[r].numpy()[base:base + D]
                if lay.pool[f] == 1 and e > s:
                    acc = acc * np.float32(1.0 / (e - s))
                o[b, col:col + D] = acc
        if sel is not None:
            return out
        res = torch.from_numpy(o)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def _check_mirror(self, tables, rf_w_off, feat_rows, feat_block, feat_owner, lay, W, mirror, feat_mirror_off):
        F = lay.num_features
        m, mo = mirror.numpy(), feat_mirror_off.tolist()
        blocks, owners, rows, w_off = feat_block.tolist(), feat_owner.tolist(), feat_rows.tolist(), rf_w_off.tolist()
        for f in range(F):
            if mo[f] < 0:
                continue
            D, i = lay.dim[f], 0
            while i < rows[f]:            # one contiguous piece per owning rank
                r, loc = self._owner_of(i, blocks[f], owners[f], W)
                n = min(rows[f] - i, blocks[f] - loc) if r < W - 1 else rows[f] - i
                base = w_off[r * F + f] + loc * D
                assert np.array_equal(m[mo[f] + i * D:mo[f] + (i + n) * D],
                                      tables.everyone[r].numpy()[base:base + n * D]), (f, i)
                i += n

    def peer_seq_gather_fwd(self, tables, rf_w_off, feat_rows, feat_block, feat_owner, lay, ids, offsets, B, W,
                            mirror=None, feat_mirror_off=None):
        if mirror is not None:
            self._check_mirror(tables, rf_w_off, feat_rows, feat_block, feat_owner, lay, W, mirror, feat_mirror_off)
        F, D = lay.num_features, lay.dim[0]
        blocks, owners = feat_block.tolist(), feat_owner.tolist()
        rows, w_off = feat_rows.tolist(), rf_w_off.tolist()
        idl, off = ids.tolist(), offsets.tolist()
        o = np.zeros((len(idl), D), dtype=np.float32)
        for f in range(F):
            for l in range(off[f * B], off[(f + 1) * B]):
                i = idl[l] if 0 <= idl[l] < rows[f] else 0
                r, loc = self._owner_of(i, blocks[f], owners[f], W)
                base = w_off[r * F + f] + loc * D
                o[l] = tables.everyone[r].numpy()[base:base + D]
        return torch.from_numpy(o)

    def peer_bucketize(self, ids, offsets, F, B, W, feat_block, feat_owner, feat_rows, rf_key_base, pooled, cap,
                       wire_key, wire_idx, counts):
        blocks, owners, rows = feat_block.tolist(), feat_owner.tolist(), feat_rows.tolist()
        kb = rf_key_base.tolist()
        idl, off = ids.tolist(), offsets.tolist()
        fill = [0] * W
        for bag in range(F * B):
            f = bag // B
            if blocks[f] <= 0:         # feature kept off the wire (small table, reduced at the source)
                continue
            for l in range(off[bag], off[bag + 1]):
                i = idl[l] if 0 <= idl[l] < rows[f] else 0
                r, loc = self._owner_of(i, blocks[f], owners[f], W)
                if fill[r] < cap:
                    wire_key[r * cap + fill[r]] = kb[r * F + f] + loc
                    wire_idx[r * cap + fill[r]] = bag if pooled else l
                fill[r] += 1
        for r in range(W):
            counts[r] = min(fill[r], cap)
        counts[W] = int(any(c > cap for c in fill))

    def peer_publish_grad(self, grad, lay, offsets, B, dst):
        g = _np(grad).copy()
        off = _np(offsets)
        for f in range(lay.num_features):
            if lay.pool[f] == 1:
                L = np.diff(off[f * B:(f + 1) * B + 1]).astype(np.float32)
                sc = np.where(L > 0, np.float32(1.0) / np.maximum(L, 1), 0).astype(np.float32)
                g[:, lay.col[f]:lay.col[f] + lay.dim[f]] *= sc[:, None]
        dst.copy_(torch.from_numpy(g))

    def peer_allreduce_mean(self, srcs, W, n, out):
        acc = srcs.everyone[0].numpy()[:n].copy()
        for r in range(1, W):
            acc = acc + srcs.everyone[r].numpy()[:n]
        out[:n] = torch.from_numpy(acc * np.float32(1.0 / W))

    def peer_barrier(self, pads, me, W, epoch):
        raise NotImplementedError("the model's barrier is a threading.Barrier (tests swap PeerBase._barrier)")

    def fused_bwd_workspace_bytes(self, lay, nnz):
        return 256

    def fused_bwd_sort_peer(self, wire_key, wire_idx, counts, me, W, cap, idx_span, lay, overflow, ws):
        keys, vals = [], []
        for r in range(W):
            c = counts.everyone[r]
            n = int(c[me])
            keys.append(wire_key.everyone[r].numpy()[me * cap:me * cap + n].copy())
            vals.append(r * idx_span + wire_idx.everyone[r].numpy()[me * cap:me * cap + n].astype(np.int64)
                        if idx_span else r * cap + np.arange(n, dtype=np.int64))      # slot mode: receive-buffer row
            if overflow is not None and int(c[W]):
                overflow |= 1
        if not hasattr(self, "_peer_sorted"):
            self._peer_sorted = {}
        self._peer_sorted[ws.data_ptr()] = (np.concatenate(keys), np.concatenate(vals))   # slot order (src-major)

    def peer_push_grad(self, recv, grad, lay, offsets, wire_idx, counts, me, W, cap, B, pooled):
        g, off = _np(grad), _np(offsets)
        D = lay.dim[0]
        for r in range(W):
            dst = recv.everyone[r].numpy()
            for j in range(int(counts[r])):
                idx = int(wire_idx[r * cap + j])
                if pooled:
                    f, b = divmod(idx, B)
                    row = g[b, lay.col[f]:lay.col[f] + D]
                    if lay.pool[f] == 1:
                        row = row * np.float32(1.0 / (off[idx + 1] - off[idx]))
                else:
                    row = g[idx, :D]
                s = (me * cap + j) * D
                dst[s:s + D] = row

    def fused_bwd_apply(self, optimizer, pooled, grad_out, weights, state, lay, offsets, nnz, B, lr, eps, grad_scale, ws,
                        **ex):
        """Two uses of the peer path are modelled: the slot-mode update (after fused_bwd_sort_peer with idx_span = 0)
        and the small tables' per-row gradient sums (optimizer = TZK_OPT_ACCUM_OUT after fused_bwd_sort)."""
        if optimizer == 100:
            return self._accum_out(pooled, grad_out, weights, state, lay, grad_scale, ws)
        assert not pooled and ws.data_ptr() in getattr(self, "_peer_sorted", {})

        class _Local:       # every "rank"'s gradient is the local receive buffer
            def __init__(self, t):
                self.everyone = {0: t}

        self.fused_bwd_apply_peer(optimizer, False, _Local(grad_out.reshape(-1)), grad_out.shape[1], weights, state, lay,
                                  B, 0, 1, nnz, nnz + 1, lr, eps, grad_scale, ws, **ex)

    def fused_bwd_apply_peer(self, optimizer, pooled, grads, ld_grad, weights, state, lay, B, me, W, cap, idx_span, lr,
                             eps, grad_scale, ws, **ex):
        from torcheasyrec_b200.kernels import FeatureLayout

        keys, vals = self._peer_sorted.pop(ws.data_ptr())
        # one "feature" per distinct table, contributions in slot order (the update's stable sort keeps that order)
        tabs = sorted({(lay.key_base[f], lay.w_off[f], lay.rows[f], lay.dim[f]) for f in range(lay.num_features)
                       if lay.rows[f] > 0})
        D = lay.dim[0]
        col_of = {}
        for f in range(lay.num_features):
            col_of[f] = lay.col[f]
        rows_out, ids_out, lens = [], [], []
        for (kb, wo, rows, dim) in tabs:
            sel = np.nonzero((keys >= kb) & (keys < kb + rows))[0]
            lens.append(len(sel))
            for s in sel:
                v = int(vals[s])
                r, idx = divmod(v, idx_span)
                if pooled:
                    f, b = divmod(idx, B)
                    gsrc = grads.everyone[r].numpy()[:B * ld_grad].reshape(B, ld_grad)[b, col_of[f]:col_of[f] + dim]
                else:
                    gsrc = grads.everyone[r].numpy()[idx * ld_grad:idx * ld_grad + dim]
                rows_out.append(gsrc.astype(np.float32))
                ids_out.append(int(keys[s]) - kb)
        tl = FeatureLayout(w_off=[t[1] for t in tabs], rows=[t[2] for t in tabs], dim=[t[3] for t in tabs],
                           col=[0] * len(tabs), pool=[0] * len(tabs), key_base=[t[0] for t in tabs],
                           total_keys=lay.total_keys, total_dim=D, arena_elems=lay.arena_elems)
        if not ids_out:
            return
        recv_g = torch.from_numpy(np.stack(rows_out))
        recv_ids = torch.tensor(ids_out, dtype=torch.int64)
        bounds = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
        self.fused_bwd(optimizer, False, recv_g, weights, state, tl, recv_ids, bounds, 1, lr, eps, grad_scale, **ex)

    # ------------------------------------------------------------------ DIN attention over jagged rows (tzk_din.cu)
    def din_attn_input_fwd(self, query, seq, offsets):
        q, k, off = _np(query), _np(seq), _np(offsets)
        N, Ds = k.shape
        B, Dq = q.shape
        qp = np.zeros((B, Ds), dtype=np.float32)
        qp[:, :Dq] = q
        seg = np.repeat(np.arange(B), np.diff(off))
        qn = qp[seg]
        return torch.from_numpy(np.concatenate([qn, k, qn - k, qn * k], axis=1).astype(np.float32))

    def din_attn_input_bwd(self, d_in, query, seq, offsets):
        g, q, k, off = _np(d_in), _np(query), _np(seq), _np(offsets)
        N, Ds = k.shape
        B, Dq = q.shape
        qp = np.zeros((B, Ds), dtype=np.float32)
        qp[:, :Dq] = q
        seg = np.repeat(np.arange(B), np.diff(off))
        g1, g2, g3, g4 = g[:, :Ds], g[:, Ds:2 * Ds], g[:, 2 * Ds:3 * Ds], g[:, 3 * Ds:]
        d_seq = (g2 - g3 + g4 * qp[seg]).astype(np.float32)
        dq = np.zeros((B, Ds), dtype=np.float32)
        np.add.at(dq, seg, (g1 + g3 + g4 * k).astype(np.float32))
        return torch.from_numpy(dq[:, :Dq].copy()), torch.from_numpy(d_seq)

    def jagged_softmax_wsum_fwd(self, scores, seq, offsets, max_len=0):
        s, k, off = _np(scores), _np(seq), _np(offsets)
        B = len(off) - 1
        probs = np.zeros(len(s), dtype=np.float32)
        out = np.zeros((B, k.shape[1]), dtype=np.float32)
        for b in range(B):
            lo, hi = off[b], off[b + 1]
            if max_len > 0:
                hi = min(hi, lo + max_len)
            if hi > lo:
                e = np.exp(s[lo:hi] - s[lo:hi].max()).astype(np.float32)
                p = (e / e.sum(dtype=np.float32)).astype(np.float32)
                probs[lo:hi] = p
                out[b] = (p[:, None] * k[lo:hi]).sum(axis=0, dtype=np.float32)
        return torch.from_numpy(probs), torch.from_numpy(out)

    def jagged_softmax_wsum_bwd(self, d_out, probs, seq, offsets, max_len=0):
        g, p, k, off = _np(d_out), _np(probs), _np(seq), _np(offsets)
        B = len(off) - 1
        d_s = np.zeros(len(p), dtype=np.float32)
        d_k = np.zeros_like(k)
        for b in range(B):
            lo, hi = off[b], off[b + 1]
            if max_len > 0:
                hi = min(hi, lo + max_len)
            if hi > lo:
                dp = k[lo:hi] @ g[b]
                t = (p[lo:hi] * dp).sum(dtype=np.float32)
                d_s[lo:hi] = p[lo:hi] * (dp - t)
                d_k[lo:hi] = p[lo:hi, None] * g[b][None, :]
        return torch.from_numpy(d_s), torch.from_numpy(d_k)

    # ------------------------------------------------------------------ small tables of the peer step (model)
    def fused_bwd_sort(self, pooled, lay, ids, offsets, B, ws):
        """Id half of the local backward: the model just remembers the ids for the gradient half."""
        if not hasattr(self, "_local_sorted"):
            self._local_sorted = {}
        self._local_sorted[ws.data_ptr()] = (_np(ids).copy(), _np(offsets).copy(), B, bool(pooled))

    def _accum_out(self, pooled, grad_out, psum, flags, lay, grad_scale, ws):
        ids, off, B, _ = self._local_sorted[ws.data_ptr()]
        g = _np(grad_out)
        P, Fl = psum.numpy(), flags.numpy()
        sums = {}
        for f in range(lay.num_features):
            if lay.rows[f] <= 0:
                continue
            D = lay.dim[f]
            for bag in range(f * B, (f + 1) * B):
                for l in range(off[bag], off[bag + 1]):
                    i = ids[l] if 0 <= ids[l] < lay.rows[f] else 0
                    if pooled:
                        row = g[bag - f * B, lay.col[f]:lay.col[f] + D].astype(np.float32)
                        sc = np.float32(grad_scale) / np.float32(off[bag + 1] - off[bag]) if lay.pool[f] == 1 \
                            else np.float32(grad_scale)
                    else:
                        row, sc = g[l, :D].astype(np.float32), np.float32(grad_scale)
                    key = lay.key_base[f] + i
                    cur = sums.get(key)
                    sums[key] = (row * sc if cur is None else cur[0] + row * sc, lay.w_off[f] + i * D, D)
        for key, (acc, o, D) in sums.items():
            P[o:o + D] = acc
            Fl[key] = 1

    def peer_small_update(self, optimizer, psum, flags, W, tabs, n_tabs, total_rows, max_dim, weights, state, lr, eps,
                          **ex):
        from torcheasyrec_b200.kernels import FeatureLayout

        rec = np.dtype([("kb", "<i8"), ("start", "<i8"), ("w_off", "<i8"), ("psum_off", "<i8"), ("key_base", "<i8"),
                        ("first", "<i4"), ("n", "<i4"), ("dim", "<i4"), ("pad", "<i4")])
        T = tabs.numpy()[:n_tabs * rec.itemsize].view(rec)
        rows_out, ids_out, lens = [], [], []
        for t in T:
            cnt = 0
            for i in range(int(t["n"])):
                key = int(t["kb"] + t["start"] + i)
                acc, any_ = np.zeros(int(t["dim"]), np.float32), False
                for r in range(W):
                    if int(flags.everyone[r][key]):
                        any_ = True
                        o = int(t["psum_off"] + (t["start"] + i) * t["dim"])
                        acc = acc + psum.everyone[r].numpy()[o:o + int(t["dim"])]
                if any_:
                    rows_out.append(acc)
                    ids_out.append(i)
                    cnt += 1
            lens.append(cnt)
        if not ids_out:
            return
        D = int(T[0]["dim"])
        assert all(int(t["dim"]) == D for t in T), "model: one dim per group"
        tl = FeatureLayout(w_off=[int(t["w_off"]) for t in T], rows=[int(t["n"]) for t in T], dim=[D] * len(T),
                           col=[0] * len(T), pool=[0] * len(T), key_base=[int(t["key_base"]) for t in T],
                           total_keys=1, total_dim=D, arena_elems=weights.numel())
        bounds = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
        self.fused_bwd(optimizer, False, torch.from_numpy(np.stack(rows_out)), weights, state, tl,
                       torch.tensor(ids_out, dtype=torch.int64), bounds, 1, lr, eps, 1.0, **ex)
