"""Summarise an `ncu --set full` capture (.ncu-rep) into profiles/: a text summary of the headline counters
and an entry in profiles/ncu_stats.json that bench.py reads for `roofline.traffic` / `fp32_issue_frac`.

    python tools/ncu_summary.py gpurun_out/r1_march.ncu-rep C2 k_march_first profiles/r01_march_first

Runs here (no GPU needed): ncu only reads the report.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = {
    "gpu__time_duration.sum": "duration_ns",
    "smsp__inst_executed.sum": "warp_instructions",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sectors.sum": "l2_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_sectors",
    "l1tex__t_bytes.sum": "l1_bytes",
    "sm__cycles_active.avg": "sm_active_cycles",
    "sm__cycles_elapsed.max": "sm_elapsed_cycles",
    "sm__inst_executed_pipe_fma.sum": "fma_pipe_warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_busy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}

UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0,
              "second": 1e9, "us": 1e3, "ms": 1e6, "ns": 1.0}


def main():
    rep, cfg, kernel, out_prefix = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    stats = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if kernel not in name and not kernel.startswith("depth"):
            continue
        cur = {"kernel_name": name.split("(")[0].split("::")[-1]}
        for k, short in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                cur[short] = v * UNIT_SCALE.get(units[i], 1.0)
        stats = cur
        break
    if not stats:
        sys.exit(f"no kernel matching {kernel} in {rep}")
    stats["dram_bytes"] = stats.get("dram_read_bytes", 0.0) + stats.get("dram_write_bytes", 0.0)
    stats["source"] = os.path.basename(out_prefix) + "_details.txt"
    details = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    with open(out_prefix + "_details.txt", "w") as f:
        f.write(details)
    path = os.path.join(ROOT, "profiles", "ncu_stats.json")
    data = {}
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
    data.setdefault(cfg, {})[kernel] = stats
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(json.dumps(stats, indent=1))


if __name__ == "__main__":
    main()
