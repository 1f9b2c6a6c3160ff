"""Turn `ncu --set full` reports into the committed evidence: a per-kernel table (markdown) and profiles/r02_traffic.json
(per-launch DRAM bytes of the dense-layer GEMM, read by bench.py's `roofline.traffic`).

    python scripts/ncu_summary.py gpurun_out/r02_step.ncu-rep gpurun_out/r02_scores.ncu-rep ... > profiles/r02_ncu_summary.md
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "ms",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor %",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "SM %",
    "dram__bytes_read.sum": "DRAM read MB",
    "dram__bytes_write.sum": "DRAM write MB",
    "lts__t_sector_hit_rate.pct": "L2 hit %",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "L2 %",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "smsp__inst_executed.sum": "warp inst (M)",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "XU %",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue %",
}
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TO_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    for vals in r[2:]:
        if len(vals) != len(hdr):
            continue
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name":
                d["name"] = v
            elif h in WANT:
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                if "bytes" in h:
                    x *= TO_MB.get(u, 1e-6)
                if h.startswith("gpu__time"):
                    x *= TO_MS.get(u, 1e-6)
                if h.startswith("smsp__inst_executed"):
                    x /= 1e6
                d[WANT[h]] = x
        yield d


def short(name):
    n = name.split("(")[0].replace("lovr::", "").replace("void ", "")
    return n if len(n) < 70 else n[:67] + "..."


def main():
    groups = defaultdict(list)
    for rep in sys.argv[1:]:
        if not os.path.exists(rep):
            continue
        for d in rows_of(rep):
            groups[(os.path.basename(rep), short(d.get("name", "?")))].append(d)
    cols = list(dict.fromkeys(WANT.values()))
    print("| capture | kernel | launches | " + " | ".join(cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    traffic = None
    for (rep, name), ds in groups.items():
        avg = {c: sum(d.get(c, 0.0) for d in ds) / len(ds) for c in cols}
        print(f"| {rep} | `{name}` | {len(ds)} | " + " | ".join(f"{avg[c]:.3g}" for c in cols) + " |")
        if name.startswith("gemm2_tcgen05_kernel") and traffic is None:
            pass
    # dense-layer GEMM traffic: average over every gemm2 launch of the step capture
    g = [d for (rep, name), ds in groups.items() if name.startswith("gemm2_tcgen05_kernel") for d in ds]
    if g:
        per_launch = sum(d.get("DRAM read MB", 0) + d.get("DRAM write MB", 0) for d in g) / len(g) * 1e6
        traffic = dict(gemm_dram_bytes_per_launch=per_launch, launches_averaged=len(g),
                       note="dram__bytes_read.sum + dram__bytes_write.sum averaged over the dense-layer GEMM launches of one "
                            "ncu --set full capture of a bench step at 128 images per native call (profiles/r02_ncu_summary.md); "
                            "scale by images per call / 128 for other chunk sizes")
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
