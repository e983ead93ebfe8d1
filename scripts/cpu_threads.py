"""CPU-baseline thread scan (run on the GPU box's host): samples/s of the oracle-port step per thread count."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    n = int(sys.argv[1])
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    os.environ.setdefault("GOMP_SPINCOUNT", "0")
    sys.path.insert(0, ROOT)
    import torch
    import bench
    torch.set_num_threads(n)
    os.cpu_count = lambda: n          # bench.cpu_step_rate sizes its pools from cpu_count
    rate, ms, cores, kind = bench.cpu_step_rate("dlrm_criteo", 8192, 3, 1, 4_000_000, "uniform")
    print(json.dumps({"threads": n, "samples_per_s": rate, "ms_per_step": ms, "backend": kind}))
else:
    for n in (8, 16, 32, 64, 128):
        if n <= (os.cpu_count() or 1):
            subprocess.run([sys.executable, __file__, str(n)])
