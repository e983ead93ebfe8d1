"""Standalone timings of the sparse kernels at DLRM-Criteo shape (full hash sizes, B=65536), CUDA events.
Usage: python scripts/bench_kernels.py            -> runs every variant in a subprocess and prints one JSON line each
       TZK_BWD_TILE=0 python scripts/bench_kernels.py one   (general backward kernels instead of the tile path)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "one":
    import torch
    from torcheasyrec_b200.example_configs import CRITEO_HASH_SIZES
    from torcheasyrec_b200.kernels import OPT_ADAGRAD, build_layout, default_kernels
    dev, B, F, D = "cuda", 65536, 26, 16
    k = default_kernels()
    lay = build_layout(CRITEO_HASH_SIZES, [D] * F, list(range(F)), [0] * F).to(dev)
    arena = torch.rand(lay.arena_elems, device=dev) * 0.01
    state = torch.zeros(lay.arena_elems, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(1)
    R = 4
    ids = [torch.cat([torch.randint(0, h, (B,), device=dev, generator=g) for h in CRITEO_HASH_SIZES]) for _ in range(R)]
    offs = torch.arange(F * B + 1, device=dev, dtype=torch.int64)
    out = torch.empty((B, F * D), device=dev)
    grad = torch.randn((B, F * D), device=dev) * 1e-3

    ITERS = int(os.environ.get("TZK_BENCH_ITERS", "20"))      # 1 = a single launch of everything (ncu captures)

    def timeit(fn, n=ITERS):
        for i in range(3 if n > 1 else 0): fn(i)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        torch.cuda.synchronize()
        for i in range(n):
            ev[i][0].record(); fn(i); ev[i][1].record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        return ts[len(ts) // 2] * 1e3
    res = {"bwd_tile": os.environ.get("TZK_BWD_TILE", "0"), "id_dist": os.environ.get("TZK_ID_DIST", "uniform")}
    if res["id_dist"] == "zipf":       # Zipf(1.05)-like skew: hot ids repeat inside a batch
        def zipf(h):
            u = torch.rand(B, device=dev, generator=g)
            return (torch.floor((h ** u) - 1).clamp_(0, h - 1)).to(torch.int64)
        ids = [torch.cat([zipf(float(h)) for h in CRITEO_HASH_SIZES]) for _ in range(R)]
    res["gather_us"] = timeit(lambda i: k.pooled_gather_fwd(arena, lay, ids[i % R], offs, B, out))
    res["fused_bwd_us"] = timeit(lambda i: k.fused_bwd(OPT_ADAGRAD, True, grad, arena, state, lay, ids[i % R], offs, B, 1e-3, 1e-8, 1.0))
    if os.environ.get("TZK_BENCH_MIN"):
        print(json.dumps(res)); sys.exit(0)
    # DLRM interaction (fwd / bwd) at the bench shape, and the narrow tower layers
    dense16 = torch.randn(B, 16, device=dev)
    sparse = torch.randn(B, F * D, device=dev)
    res["interact_fwd_us"] = timeit(lambda i: k.dot_interact_fwd(dense16, sparse, F, D, True, True, 4, 1))
    d_out = torch.randn(B, 784, device=dev)
    res["interact_bwd_us"] = timeit(lambda i: k.dot_interact_bwd(dense16, sparse, d_out, F, D, True, True, 1))
    for (K, N) in ((13, 64), (64, 16), (64, 32), (32, 1)):
        x = torch.randn(B, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
        y = k.small_linear_fwd(x, w, b, True); dy = torch.randn(B, N, device=dev)
        res[f"lin{K}x{N}_fwd_us"] = timeit(lambda i: k.small_linear_fwd(x, w, b, True))
        res[f"lin{K}x{N}_bwd_us"] = timeit(lambda i: k.small_linear_bwd(x, w, y, dy, True, True, True))
    print(json.dumps(res))
else:
    for tile, dist in (("0", "uniform"), ("1", "uniform"), ("0", "zipf")):
        env = dict(os.environ, TZK_BWD_TILE=tile, TZK_ID_DIST=dist)
        subprocess.run([sys.executable, __file__, "one"], env=env)
