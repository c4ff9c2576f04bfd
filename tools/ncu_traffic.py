"""profiles/r02_traffic.json from an ncu CSV of `bench.py` (metrics dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum): DRAM bytes of ONE assembly step = the last launch of each assembly kernel.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python bench.py --steps 2 --warmup 3 --no-solve --cfg2-size 0 --no-cpu-baseline --size S
    python tools/ncu_traffic.py gpurun_out/traffic.csv elasticity S staged [more csv workload size mode ...]
"""
import csv
import json
import os
import sys

ASSEMBLY = ("element_dmma_kernel", "element_kernel", "element_nh_dmma_kernel", "hex27_kernel", "gather_residual_kernel",
            "apply_bc_vec_kernel", "gather_csr_kernel", "staged_assembly_kernel", "fused_assembly_kernel", "fused_dmma_kernel")


def capture(path, workload, size, mode):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].strip('"').isdigit()]
    last = {}
    for r in rows:
        name = next((k for k in ASSEMBLY if k in r[4]), None)
        if name is None:
            continue
        metric, value = r[-3], float(r[-1].replace(",", ""))
        last.setdefault(name, {})[metric] = value        # later launches overwrite earlier ones
    kernels = {k: {"dram_read": v.get("dram__bytes_read.sum", 0.0), "dram_write": v.get("dram__bytes_write.sum", 0.0),
                   "ncu_time_ms": v.get("gpu__time_duration.sum", 0.0) / 1e6} for k, v in last.items()}
    total = sum(v["dram_read"] + v["dram_write"] for v in kernels.values())
    return {"workload": workload, "size": int(size), "mode": mode, "assembly_bytes_per_step": total, "kernels": kernels,
            "source": os.path.basename(path), "note": "ncu serialises and cold-starts every launch: bytes are per launch, times are not bench values"}


if __name__ == "__main__":
    args = sys.argv[1:]
    caps = [capture(*args[i:i + 4]) for i in range(0, len(args), 4)]
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
    json.dump({"captures": caps}, open(out, "w"), indent=1)
    for c in caps:
        print(c["workload"], c["size"], c["mode"], f"{c['assembly_bytes_per_step'] / 1e9:.3f} GB", {k: round((v['dram_read'] + v['dram_write']) / 1e9, 3) for k, v in c["kernels"].items()})
