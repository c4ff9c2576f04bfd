"""Turn ncu exports into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.md
    ncu -i X.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/summarize_ncu.py raw /tmp/raw.csv > profiles/r01_kernels.md
"""
import csv
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
]


def short(name):
    name = name.replace("void ", "").replace("femb200::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.split("(")[0][:70]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= i_val:
            continue
        v = float(r[i_val].replace(",", ""))
        v = {"ns": v * 1e-6, "us": v * 1e-3, "ms": v, "nsecond": v * 1e-6, "usecond": v * 1e-3, "msecond": v}.get(r[i_unit], v * 1e-6)
        k = short(r[i_name])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print("| kernel | launches | total ms | mean ms | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.3f} | {t / n:.4f} | {100 * t / total:.1f}% |")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        if name in seen:
            continue
        seen.add(name)
        print(f"### `{name}`\n")
        print("| metric | value |")
        print("|---|---|")
        for m, label in METRICS:
            if m in idx:
                print(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
        st = sorted(((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stall), reverse=True)[:5]
        print("| top warp stall reasons (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |")
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
