#!/usr/bin/env python
"""Turns the ncu artefacts brought back in gpurun_out/ into the committed summaries
under profiles/ (run in the build container, where ncu can read .ncu-rep files):

    python profiles/summarize.py r01

writes profiles/<tag>_launches.md (share of each kernel in the step),
profiles/<tag>_kernels.md (key metrics of the --set full captures) and updates
profiles/mac_traffic.json (dram bytes per MAC launch, read by bench.py)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    # ---- launch list
    src = os.path.join(G, f"{tag}_launches.csv")
    if not os.path.exists(src):
        src = os.path.join(P, f"{tag}_launches.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        key = (r[4].split("(")[0].replace("void ", ""), r[8])
        agg.setdefault(key, []).append(float(r[-1]) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)\n\n"
                "Command: `python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e` (santalucia, 1024 streams;\n"
                "r01: `--blocks-per-step 1`, r01tt: 4 blocks per step, r01f: 8 blocks per step = bench.py's default).\n"
                "Rows with 128 streams in the grid are the 8 chunks of the one warm-up call through the host-staging\n"
                "path (fcv_batch_process); the 1024-stream rows are the device-resident loop that `value` times.\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's\n"
                "`kernel_ms_per_step`, not absolutes.\n\n| kernel | grid | launches | mean us | share of listed time |\n|---|---|---|---|---|\n")
        for (k, grid), v in agg.items():
            f.write(f"| {k} | {grid} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / tot:.3f} |\n")
    # ---- full captures
    traffic_path = os.path.join(P, "mac_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    with open(os.path.join(P, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` captures\n")
        keys = dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a)   # e.g. r01_mactt4.ncu-rep=santalucia:1024:T4
        for rep in [f"{tag}_mac.ncu-rep", f"{tag}_fft.ncu-rep"] + [k for k in keys if k not in (f"{tag}_mac.ncu-rep",)]:
            path = os.path.join(G, rep)
            if not os.path.exists(path):
                continue
            hdr, units, data = raw(path)
            for r in data:
                name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
                f.write(f"\n## {name}  ({rep})\n\n| metric | value | unit |\n|---|---|---|\n")
                for w in WANT:
                    if w in hdr:
                        f.write(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
                if name.startswith("mac_"):
                    i, j = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                    b = to_bytes(r[i], units[i]) + to_bytes(r[j], units[j])
                    grid = int(r[hdr.index("launch__grid_size")].replace(",", ""))
                    streams = grid // 32 * 4  # grid = 32 bin tiles x ceil(streams/4) for fragm 8192
                    traffic[keys.get(rep, f"santalucia:{streams}")] = {
                        "dram_bytes_per_launch": b, "source": f"profiles/{tag}_kernels.md ({rep})",
                        "duration_us_under_ncu": float(r[hdr.index('gpu__time_duration.sum')].replace(',', ''))}
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print(open(os.path.join(P, f"{tag}_launches.md")).read())
    print(open(os.path.join(P, f"{tag}_kernels.md")).read())
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
