"""Aggregate the per-instruction stall samples of an ncu report (--page source --csv) into the regions between
barriers, and list the hottest instructions.   python tools/ncu_regions.py X.ncu-rep [top]"""
import csv
import subprocess
import sys


def main(rep, top=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    k0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[k0]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[k0 + 1:] if len(r) == len(hdr)]
    tot = sum(int(r[ci['# Samples']]) for r in data)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    print(rows[0][1] if rows[0] else '', '| samples', tot, '| instructions', len(data))
    start, cur = 0, [0, 0, {}]
    for i, r in enumerate(data):
        cur[0] += int(r[ci['# Samples']])
        cur[1] += int(r[ci['Instructions Executed']])
        for k in stalls:
            cur[2][k] = cur[2].get(k, 0) + int(r[ci[k]])
        if 'BAR.SYNC' in r[1] or i == len(data) - 1:
            t = sorted(cur[2].items(), key=lambda kv: -kv[1])[:4]
            print(f"sass {start:5d}-{i:5d}  samples {cur[0]:7d} ({100 * cur[0] / max(tot, 1):5.1f}%)  warp-inst {cur[1]:11d}  " +
                  ", ".join(f"{k[6:]} {v}" for k, v in t))
            start, cur = i + 1, [0, 0, {}]
    if top:
        print("hottest instructions:")
        for i, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][ci['# Samples']]))[:top]:
            t = sorted(((k, int(r[ci[k]])) for k in stalls), key=lambda kv: -kv[1])[:2]
            print(f"  {i:5d} {int(r[ci['# Samples']]):7d}  {r[1].strip()[:70]:70s} " + ", ".join(f"{k[6:]} {v}" for k, v in t))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
