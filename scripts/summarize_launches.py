"""Per-step kernel table from an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.

    python scripts/summarize_launches.py gpurun_out/launches.csv [--md]

Finds one steady-state graph replay (kernels between two consecutive `scan_tile_sums` launches, the first own kernel
of a step) and prints every launch with its duration plus totals per group (own sparse path / own tower kernels /
library GEMMs / torch element-wise)."""
import csv
import sys
from collections import Counter


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    return [(r[kn], float(r[mv].replace(",", "")) / 1e3) for r in rows[hi + 1:] if len(r) > mv and r[mv]]


def group(name):
    own_sparse = ("scan_tile", "pooled_gather", "seq_gather", "linearize", "run_update", "long_chunk", "long_combine",
                  "fused_apply", "find_long_runs", "tile_update", "carry_combine", "zero_counters", "bucketize", "bag_grad",
                  "permute_", "col_gather", "jagged", "fm_", "dot_interact", "peer_", "small_table_update", "din_",
                  "softmax_wsum")
    own_tower = ("small_linear", "bce_", "bias_act", "act_bwd_colsum", "colsum_final", "tower_tail")
    own_gemm = ("gemm3x_kernel", "wgrad3x_kernel", "wgrad_reduce", "split_w_kernel")
    if "DeviceRadixSort" in name:
        return "radix sort (CUB, inside tzk_fused_bwd)"
    if any(k in name for k in own_sparse):
        return "own: sparse path + interaction"
    if any(k in name for k in own_tower):
        return "own: tower / loss kernels"
    if any(k in name for k in own_gemm):
        return "own: tcgen05 3xTF32 GEMMs of the wide tower layer"
    if "gemm" in name.lower() or "inf_patching" in name or "splitK" in name or "cutlass" in name:
        return "library GEMM (cuBLASLt BF16x9 + its inf/nan scans)"
    return "torch element-wise / optimizer"


def main():
    data = load(sys.argv[1])
    idx = [i for i, (n, _) in enumerate(data) if "scan_tile_sums" in n]
    gaps = [idx[k + 1] - idx[k] for k in range(len(idx) - 1)]
    L = Counter(gaps).most_common(1)[0][0]
    k = next(k for k in range(2, len(gaps)) if gaps[k] == L)
    seg = data[idx[k]:idx[k + 1]]
    tot = sum(v for _, v in seg)
    md = "--md" in sys.argv
    print(f"{len(seg)} kernels per step, {tot:.0f} us serialised under ncu")
    by = Counter()
    for n, v in seg:
        by[group(n)] += v
    for g, v in by.most_common():
        print(f"| {g} | {v:.0f} | {100 * v / tot:.0f} % |" if md else f"  {g:60s} {v:8.1f} us  {100 * v / tot:4.1f} %")
    print()
    for i, (n, v) in enumerate(seg):
        short = n.replace("void ", "").replace("<unnamed>::", "")[:90]
        print(f"| {i} | `{short}` | {v:.1f} |" if md else f"{i:3d} {short:92s} {v:7.1f}")


if __name__ == "__main__":
    main()
