"""Turn `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/ (rNN_ncu_full_hot_kernels.txt):
duration, DRAM MB read / written, tensor-pipe %, DRAM %, achieved warps %, registers, grid size.

    ncu -i gpurun_out/prof_targets.ncu-rep --page raw --csv > gpurun_out/prof_targets_raw.csv     # runs here, no GPU needed
    python scripts/ncu_raw_table.py gpurun_out/prof_targets_raw.csv > profiles/rNN_ncu_full_hot_kernels.txt

Handles both CSV layouts ncu emits: wide (one row per launch, one column per metric, a units row under the header) and long
(one row per launch and metric with "Metric Name" / "Metric Unit" / "Metric Value" columns)."""
import csv
import re
import sys

WANT = {
    "us": ("gpu__time_duration.sum",),
    "rd": ("dram__bytes_read.sum",),
    "wr": ("dram__bytes_write.sum",),
    "tensor": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
               "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
    "dram": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    "warps": ("sm__warps_active.avg.pct_of_peak_sustained_active",),
    "regs": ("launch__registers_per_thread",),
    "grid": ("launch__grid_size",),
}
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "B": 1e-6, "KB": 1e-3, "MB": 1.0, "GB": 1e3}


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def short(name):
    clean = name.replace("(anonymous namespace)::", "")
    m = re.search(r"([A-Za-z_][A-Za-z0-9_:]*)(<[^()]*>)?\(", clean)
    return ((m.group(1) + (m.group(2) or "")) if m else clean)[:78]


def main(path):
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    header = rows[0]
    launches = []                                   # [(kernel name, {metric: (value, unit)})]
    if "Metric Name" in header:                     # long layout
        col = {h: i for i, h in enumerate(header)}
        cur = {}
        for r in rows[1:]:
            key = (r[col["ID"]], r[col["Kernel Name"]])
            cur.setdefault(key, {})[r[col["Metric Name"]]] = (num(r[col["Metric Value"]]), r[col["Metric Unit"]])
        launches = [(k[1], v) for k, v in cur.items()]
    else:                                           # wide layout: header, units, then one row per launch
        units = rows[1]
        kcol = header.index("Kernel Name")
        for r in rows[2:]:
            if len(r) != len(header):
                continue
            launches.append((r[kcol], {h: (num(v), u) for h, v, u in zip(header, r, units)}))
    print("%-80s %8s %8s %8s %8s %7s %7s %6s %8s" % ("kernel", "us", "MB rd", "MB wr", "tensor%", "dram%", "warps%", "regs", "grid"))
    summary = {}
    for name, m in launches:
        def get(key, scale=None):
            for cand in WANT[key]:
                if cand in m and m[cand][0] is not None:
                    v, u = m[cand]
                    return v * (scale.get(u, 1.0) if scale else 1.0)
            return float("nan")
        print("%-80s %8.1f %8.1f %8.1f %8.1f %7.1f %7.1f %6.0f %8.0f" % (short(name), get("us", TIME), get("rd", BYTES),
              get("wr", BYTES), get("tensor"), get("dram"), get("warps"), get("regs"), get("grid")))
        key = short(name)
        k = 2
        while key in summary:                      # several launches of one kernel (different shapes): name, name#2, ...
            key = "%s#%d" % (short(name), k)
            k += 1
        summary[key] = {"us": round(get("us", TIME), 2), "dram_read_mb": round(get("rd", BYTES), 2),
                        "dram_write_mb": round(get("wr", BYTES), 2), "tensor_pct": round(get("tensor"), 2),
                        "dram_pct": round(get("dram"), 2), "grid": int(get("grid"))}
    return summary


if __name__ == "__main__":
    # python scripts/ncu_raw_table.py raw.csv [--json profiles/ncu_hot_kernels.json --commit HASH --order "name; name; ..."]
    import argparse
    import datetime
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--json", default=None, help="also write the machine-readable summary bench.py reads")
    ap.add_argument("--commit", default="unknown", help="git commit of the profiled code")
    ap.add_argument("--gemm", default=None, help="summary key of the dominant GEMM launch (default: first gemm_* entry)")
    ap.add_argument("--gemm-what", default="dominant GEMM launch")
    a = ap.parse_args()
    summ = main(a.csv)
    if a.json:
        summ = {k: v for k, v in summ.items() if not k.startswith("at::")}      # L2-flush fills between the profiled launches
        gk = a.gemm or next((k for k in summ if "gemm" in k), None)
        out = {"commit": a.commit, "when": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ"),
               "how": "ncu --set full --clock-control none, L2 flushed before each profiled launch (scripts/gpu_ncu_targets.py)",
               "kernels": summ}
        if gk:
            g = summ[gk]
            out["dominant_gemm"] = {"kernel": gk, "what": a.gemm_what,
                                    "dram_bytes": round((g["dram_read_mb"] + g["dram_write_mb"]) * 1e6),
                                    "note": "%.1f MB read + %.1f MB written to DRAM inside the capture window, %.1f us, tensor pipe "
                                            "%.1f %%" % (g["dram_read_mb"], g["dram_write_mb"], g["us"], g["tensor_pct"])}
        with open(a.json, "w") as fh:
            json.dump(out, fh, indent=1)
