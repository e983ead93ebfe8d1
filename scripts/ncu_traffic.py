"""One `ncu --set full` capture of the gather and of every kernel of the fused backward (sort included) on the headline
workload -> profiles/ncu_traffic.json (dram read + write bytes per launch, per kernel and summed) + the raw CSV and the
.ncu-rep next to it.  bench.py reads the JSON for `roofline.traffic` and ignores it when the kernel sources changed.

    gpurun --timeout 900 -- 'python scripts/ncu_traffic.py r2'      # writes gpurun_out/ncu_traffic/*, then copy to profiles/
"""
import csv
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = "regex:linearize|RadixSort|fused_apply|find_long_runs|run_update|long_chunk|long_combine|small_table|pooled_gather_fwd|zero_counters|tile_update|carry_combine"


def source_sha16():
    h = hashlib.sha256()
    for fn in ("tzk_bwd.cu", "tzk_gather.cu", "tzk_common.cuh"):
        h.update(open(os.path.join(ROOT, "torcheasyrec_b200", "csrc", fn), "rb").read())
    return h.hexdigest()[:16]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    out = os.path.join(ROOT, "gpurun_out", "ncu_traffic")
    os.makedirs(out, exist_ok=True)
    rep = os.path.join(out, f"{tag}_fused_bwd_gather")
    reps = 2
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", KERNELS, "-f", "-o", rep,
           sys.executable, os.path.join(ROOT, "scripts", "ncu_target.py"), str(reps)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    open(os.path.join(out, f"{tag}_ncu.log"), "w").write(r.stdout[-20000:] + "\n---- stderr ----\n" + r.stderr[-20000:])
    if r.returncode != 0 or not os.path.exists(rep + ".ncu-rep"):
        print("ncu failed", r.returncode, r.stderr[-2000:])
        return 1
    raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv", "--print-units", "base"],
                         capture_output=True, text=True, check=True).stdout
    open(os.path.join(out, f"{tag}_raw.csv"), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    need = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"]
    for n in need:
        assert n in col, f"{n} missing from the ncu raw page"
    launches = []
    for row in rows[2:]:
        if len(row) < len(hdr):
            continue
        name = row[col["Kernel Name"]]
        f = lambda k: float(row[col[k]].replace(",", "") or 0)
        launches.append({"kernel": name.split("(")[0][:120], "dram_read": f("dram__bytes_read.sum"),
                         "dram_write": f("dram__bytes_write.sum"), "ns": f("gpu__time_duration.sum")})
    # the last repetition of the target = the last 1/reps of the launch list
    per_rep = len(launches) // reps
    last = launches[-per_rep:]
    gather = [x for x in last if "pooled_gather_fwd" in x["kernel"]]
    bwd = [x for x in last if "pooled_gather_fwd" not in x["kernel"]]
    doc = {"source_sha16": source_sha16(), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "how": "ncu --set full --clock-control none, stand-alone launches of scripts/ncu_target.py (dlrm_criteo, "
                  "B=65536, uniform ids), last of 2 repetitions; dram__bytes_read.sum + dram__bytes_write.sum",
           "kernels": [x["kernel"].split("<")[0].split("::")[-1] for x in bwd],
           "fused_bwd_bytes": sum(x["dram_read"] + x["dram_write"] for x in bwd),
           "pooled_gather_fwd_bytes": sum(x["dram_read"] + x["dram_write"] for x in gather),
           "per_kernel": last}
    json.dump(doc, open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in doc.items() if k != "per_kernel"}))
    for x in last:
        print(f"{x['kernel'][:70]:<70} {x['ns'] / 1e3:8.1f} us  rd {x['dram_read'] / 1e6:8.1f} MB  wr {x['dram_write'] / 1e6:8.1f} MB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
