"""The peer-memory sparse step (csrc/tzk_peer.cu + the peer mode of csrc/tzk_bwd.cu) on ONE GPU: W virtual ranks are
threads of this process, their "symmetric" buffers are ordinary allocations of the same device and the device barrier
is "drain the GPU, then meet" — everything else is the real thing: CUDA kernels through the C-ABI, `PeerState`, side
streams.  (Two real GPUs over NVLink: tests/test_distributed_gpu.py.)  Checked against the unsharded CUDA collection."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

pytestmark = pytest.mark.gpu


def _helpers():
    import test_peer_exchange_model as M

    return M


def _seed_groups(cfgs, plan, W, pooled, full, spec, alpha):
    from torcheasyrec_b200.distributed import TABLE_WISE, _DimGroup
    from torcheasyrec_b200.embedding_modules import output_names_by_table

    names = output_names_by_table(cfgs)
    groups = []
    for r in range(W):
        g = _DimGroup(cfgs, plan, r, W, torch.device("cuda"), pooled, names)
        g.static_alpha = alpha
        if spec is not None:
            g.local.set_optimizer(spec)
        for t, c in enumerate(cfgs):
            n = g.local._table_rows[t]
            if n:
                start = 0 if plan[c.name].kind == TABLE_WISE else r * plan[c.name].block
                g.local.set_table_weight(t, full.table_weight(t)[start:start + n])
        groups.append(g)
    return groups


def _gathered(groups, plan, cfgs, full, t):
    from torcheasyrec_b200.distributed import TABLE_WISE

    sh = plan[cfgs[t].name]
    got = torch.zeros_like(full.table_weight(t))
    for r, g in enumerate(groups):
        n = g.local._table_rows[t]
        if n:
            start = 0 if sh.kind == TABLE_WISE else r * sh.block
            got[start:start + n] = g.local.table_weight(t)
    return got


@pytest.mark.parametrize("W,multi_hot,B", [(2, False, 64), (4, False, 1000), (8, False, 4099), (3, True, 257)])
@pytest.mark.parametrize("opt,bwd", [("adagrad", "push"), ("adam", "push"), ("adagrad", "pull")])
def test_peer_pooled_step_on_one_gpu(kernels, monkeypatch, W, multi_hot, B, opt, bwd):
    """bwd: gradient transport — "push" (sources write their slices into the owners' receive buffers, wire order) or
    "pull" (the owners' update kernels read the slices from the sources' published gradients)."""
    monkeypatch.setenv("TZK_PEER_BWD", bwd)
    M = _helpers()
    from torcheasyrec_b200 import peer_exchange
    from torcheasyrec_b200.distributed import TABLE_WISE, make_plan
    from torcheasyrec_b200.embedding_modules import EmbeddingBagCollection, SparseOptimizerSpec

    torch.manual_seed(0)
    rng = np.random.default_rng(W * 31 + B)
    cfgs = M._pooled_configs()
    D = 16
    plan = make_plan(cfgs, W, "row_wise", {"t_tw": [TABLE_WISE], "t_tiny": [TABLE_WISE]})
    spec = SparseOptimizerSpec.from_name(opt, lr=0.05)
    full = EmbeddingBagCollection(cfgs, device="cuda")
    full.set_optimizer(spec)
    F = len(full.feature_names())
    feat_rows = [cfgs[t].num_embeddings for t in full._feat_table]
    groups = _seed_groups(cfgs, plan, W, True, full, spec, 2.5)
    batches = [M._bags(rng, F, B, feat_rows, multi_hot) for _ in range(W)]
    ids, offs = [b[0].cuda() for b in batches], [b[1].cuda() for b in batches]
    grads = [torch.from_numpy(rng.standard_normal((B, F * D)).astype(np.float32)).cuda() for _ in range(W)]
    registry, outs = {}, [[None, None] for _ in range(W)]
    budget = [B * (4 if multi_hot else 1)] * F

    def body(r, tbar):
        torch.cuda.set_device(0)

        class St(M._sim_mixin(registry, tbar, "gpu", "cuda"), peer_exchange.PeerState):
            pass

        st = St(groups[r], plan, None, B, budget)
        for step in range(2):
            outs[r][step] = st.gather(ids[r], offs[r])
            st.prep(ids[r], offs[r])
            st.backward(grads[r], offs[r])
            torch.cuda.synchronize()

    M._run_ranks(W, body)
    assert all(int(g.overflow.item()) == 0 for g in groups)
    k = kernels
    cat = M._cat_key_major([i.cpu() for i in ids], [o.cpu() for o in offs], F, B, W)
    cat_ids, cat_off = cat[0].cuda(), cat[1].cuda()
    cat_grad = torch.cat(grads) / W
    for step in range(2):
        for r in range(W):      # forward: the unsharded gather's bits (step 0), after one update within fp32 round-off
            want = k.pooled_gather_fwd(full.weights.data, full.layout, ids[r], offs[r], B)
            if step == 0:
                assert torch.equal(outs[r][0], want)
            else:
                torch.testing.assert_close(outs[r][1], want, rtol=2e-5, atol=1e-6)
        k.fused_bwd(spec.kind, True, cat_grad, full.weights.data, full.opt_state, full.layout, cat_ids, cat_off, B * W,
                    spec.lr, spec.eps, 1.0, **full.opt_extras())
    for t, c in enumerate(cfgs):
        torch.testing.assert_close(_gathered(groups, plan, cfgs, full, t), full.table_weight(t), rtol=5e-5, atol=1e-6,
                                   msg=lambda m, c=c: f"{c.name}: {m}")


@pytest.mark.parametrize("W", [2, 5])
def test_peer_sequence_step_on_one_gpu(kernels, W):
    M = _helpers()
    from torcheasyrec_b200 import peer_exchange
    from torcheasyrec_b200.distributed import TABLE_WISE, make_plan
    from torcheasyrec_b200.embedding_modules import EmbeddingCollection, EmbeddingConfig, SparseOptimizerSpec

    rng = np.random.default_rng(11 + W)
    mk = lambda n, rows, feats: EmbeddingConfig(num_embeddings=rows, embedding_dim=16, name=n, feature_names=feats)
    cfgs = [mk("q", 50, ["q_id"]), mk("s1", 21100, ["seq_a"]), mk("s2", 40, ["seq_b"])]
    B, D, max_len = 300, 16, 20
    plan = make_plan(cfgs, W, "row_wise", {"s2": [TABLE_WISE]})
    spec = SparseOptimizerSpec.from_name("adagrad", lr=0.1)
    full = EmbeddingCollection(cfgs, device="cuda")
    full.set_optimizer(spec)
    F = len(full.feature_names())
    feat_rows = [cfgs[t].num_embeddings for t in full._feat_table]
    batches = []
    for _ in range(W):
        lens = np.concatenate([np.ones(B, np.int64), rng.integers(0, max_len + 1, B), rng.integers(0, max_len + 1, B)])
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        idv = np.concatenate([rng.integers(0, feat_rows[b // B], lens[b]) for b in range(F * B)]).astype(np.int64)
        batches.append((torch.from_numpy(idv).cuda(), torch.from_numpy(off).cuda()))
    grads = [torch.from_numpy(rng.standard_normal((b[0].numel(), D)).astype(np.float32)).cuda() for b in batches]
    groups = _seed_groups(cfgs, plan, W, False, full, spec, 2.0)
    registry, outs = {}, [None] * W

    def body(r, tbar):
        torch.cuda.set_device(0)

        class St(M._sim_mixin(registry, tbar, "gpuseq", "cuda"), peer_exchange.PeerState):
            pass

        st = St(groups[r], plan, None, B, [B, B * max_len, B * max_len])
        outs[r] = st.gather(*batches[r])
        st.prep(*batches[r])
        st.backward(grads[r], batches[r][1])
        torch.cuda.synchronize()

    M._run_ranks(W, body)
    assert all(int(g.overflow.item()) == 0 for g in groups)
    k = kernels
    for r in range(W):
        assert torch.equal(outs[r], k.seq_gather_fwd(full.weights.data, full.layout, batches[r][0], batches[r][1], B))
    ids, offs = [b[0].cpu() for b in batches], [b[1].cpu() for b in batches]
    cat_ids, cat_off = M._cat_key_major(ids, offs, F, B, W)
    rows = []
    for f in range(F):
        for r in range(W):
            o = offs[r].numpy()
            rows.append(grads[r][o[f * B]:o[(f + 1) * B]])
    k.fused_bwd(spec.kind, False, torch.cat(rows) / W, full.weights.data, full.opt_state, full.layout, cat_ids.cuda(),
                cat_off.cuda(), B * W, spec.lr, spec.eps, 1.0)
    for t, c in enumerate(cfgs):
        torch.testing.assert_close(_gathered(groups, plan, cfgs, full, t), full.table_weight(t), rtol=5e-5, atol=1e-6,
                                   msg=lambda m, c=c: f"{c.name}: {m}")


def test_peer_overflow_flag_and_dense_sync_on_one_gpu(kernels):
    M = _helpers()
    from torcheasyrec_b200 import peer_exchange
    from torcheasyrec_b200.distributed import make_plan
    from torcheasyrec_b200.embedding_modules import EmbeddingBagConfig, SparseOptimizerSpec, EmbeddingBagCollection

    W, B = 3, 64
    cfgs = [EmbeddingBagConfig(num_embeddings=300, embedding_dim=16, name="t", feature_names=["a"]),
            EmbeddingBagConfig(num_embeddings=90, embedding_dim=16, name="u", feature_names=["b"])]
    plan = make_plan(cfgs, W, "row_wise")
    spec = SparseOptimizerSpec.from_name("adagrad", lr=0.05)
    full = EmbeddingBagCollection(cfgs, device="cuda")
    groups = _seed_groups(cfgs, plan, W, True, full, spec, 1.0)
    before = [g.local.weights.data.clone() for g in groups]
    ids = torch.cat([torch.arange(B) % 7, torch.arange(B) % 5]).to(torch.int64).cuda()     # all in rank 0's blocks
    off = torch.arange(2 * B + 1, dtype=torch.int64).cuda()
    vals = [torch.randn(1000, device="cuda") for _ in range(W)]
    params = [[torch.nn.Parameter(torch.zeros(1000, device="cuda"))] for _ in range(W)]
    registry = {}

    def body(r, tbar):
        torch.cuda.set_device(0)

        class St(M._sim_mixin(registry, tbar, "ovf", "cuda"), peer_exchange.PeerState):
            pass

        class Sync(M._sim_mixin(registry, tbar, "dense", "cuda"), peer_exchange.PeerDenseGradSync):
            pass

        st = St(groups[r], plan, None, B)
        st.gather(ids, off)
        st.prep(ids, off)
        st.backward(torch.ones(B, 32, device="cuda"), off)
        s = Sync(params[r], None, world=W, rank=r)
        s.zero()
        params[r][0].grad.add_(vals[r])
        s.sync()
        torch.cuda.synchronize()

    M._run_ranks(W, body)
    assert all(int(g.overflow.item()) == 1 for g in groups)
    for r in (1, 2):
        assert torch.equal(groups[r].local.weights.data, before[r])
    want = ((vals[0] + vals[1]) + vals[2]) * np.float32(1.0 / 3)
    for r in range(W):
        assert torch.equal(params[r][0].grad, want)
