"""Turn the ncu captures of a GPU round (gpurun_out/<tag>_cfg4_<kernel>.ncu-rep, <tag>_launches_cfg4.csv)
into the committed evidence under profiles/: per-kernel raw-page CSV, a summary JSON (time, DRAM bytes,
throughput, issue / pipe utilisation, occupancy, registers), the launch-list shares of one composite,
and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).

    python tools/ncu_summary.py r02i
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TRACE_NAMES = {"warp_tiles": "K1t_warp_tiles", "blur_h_list": "K3_gauss_blur", "blur_v_list": "K3_gauss_blur",
               "multiband_collapse": "K4_multiband_collapse", "pyramid_reduce_list": "K3a_pyramid_reduce",
               "pack_rgbx": "K1p_pack_rgbx", "pack_rgbx_batch": "K1p_pack_rgbx", "seam_candidates": "K0_seam_plan"}
WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__inst_executed.sum.pct_of_peak_sustained_elapsed": "issue_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "smsp__inst_executed.sum": "warp_instructions",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def raw_page(rep):
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(text.splitlines()))


def main():
    summary, traffic = {}, {}
    for name in sorted(os.listdir(OUT)):
        m = re.fullmatch(rf"{TAG}_cfg4_(\w+)\.ncu-rep", name)
        if not m:
            continue
        kernel = m.group(1)
        rows = raw_page(os.path.join(OUT, name))
        if len(rows) < 3:
            continue
        with open(os.path.join(PROF, f"{TAG}_ncu_full_cfg4_{kernel}.csv"), "w", newline="") as fh:
            csv.writer(fh).writerows(rows)
        head, units, vals = rows[0], rows[1], rows[2]
        rec = {"kernel_name": vals[head.index("Kernel Name")] if "Kernel Name" in head else kernel}
        for h, u, v in zip(head, units, vals):
            if h in WANT and v:
                x = float(v.replace(",", ""))
                rec[WANT[h]] = x * UNIT.get(u, 1.0) if WANT[h] in ("time", "dram_read", "dram_write") else x
        rec["dram_bytes"] = rec.get("dram_read", 0.0) + rec.get("dram_write", 0.0)
        rec["dram_GBps"] = rec["dram_bytes"] / max(rec.get("time", 1e-9), 1e-12) / 1e9
        summary[kernel] = rec
        trace = TRACE_NAMES.get(kernel)
        if trace:
            traffic[trace] = traffic.get(trace, 0.0) + rec["dram_bytes"]
    launches = os.path.join(OUT, f"{TAG}_launches_cfg4.csv")
    shares = {}
    if os.path.exists(launches):
        with open(launches) as fh:
            text = fh.read()
        with open(os.path.join(PROF, f"{TAG}_ncu_launches_cfg4.csv"), "w") as fh:
            fh.write(text)
        rows = [r for r in csv.reader(text.splitlines()) if len(r) > 5]
        head = next(r for r in rows if "Kernel Name" in r)
        ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
        per = {}
        for r in rows[rows.index(head) + 1:]:
            try:
                t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
            except ValueError:
                continue
            short = re.sub(r"\(.*", "", r[ki]).replace("void p360::", "").strip()
            per.setdefault(short, [0.0, 0])
            per[short][0] += t
            per[short][1] += 1
        total = sum(v[0] for v in per.values())
        shares = {k: {"ms": round(v[0], 4), "launches": v[1], "share": round(v[0] / total, 4)}
                  for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])}
    with open(os.path.join(PROF, f"{TAG}_ncu_summary_cfg4.json"), "w") as fh:
        json.dump({"tag": TAG, "command": "python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs "
                   "(cfg4, 1 x B200); ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 1 -c 1",
                   "kernels": summary,
                   "launch_list": {"note": "ncu --metrics gpu__time_duration.sum: every launch of the same command "
                                           "(cold-cache, serialised; warm-up + 1 timed + e2e steps): compare SHARES",
                                   "per_kernel": shares}}, fh, indent=1)
    path = os.path.join(PROF, "ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table["cfg4"] = {k: round(v) for k, v in traffic.items()}
    table["_source"] = f"profiles/{TAG}_ncu_summary_cfg4.json (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(table, open(path, "w"), indent=1)
    for k, r in summary.items():
        print(f"{k:22s} {r.get('time', 0) * 1e3:7.3f} ms  DRAM {r['dram_bytes'] / 1e9:6.2f} GB ({r['dram_GBps']:6.0f} GB/s, "
              f"{r.get('dram_pct_of_peak', 0):4.1f} %)  issue {r.get('issue_pct_of_peak', 0):4.1f} %  occ {r.get('achieved_occupancy_pct', 0):4.1f} %  "
              f"regs {int(r.get('registers', 0))}")
    for k, v in list(shares.items())[:14]:
        print(f"   {k:40s} {v}")


if __name__ == "__main__":
    main()
