"""Summarise an ncu --set full report (CSV from `ncu -i X.ncu-rep --page raw --csv`) into a small JSON:
per kernel (first capture of each name): duration, DRAM bytes read/written, throughput percentages, registers, occupancy.
usage: python scripts/ncu_summarise.py out.json report1.ncu-rep [report2.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__grid_size": "grid",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "ms ": 1e3}


def main():
    out = {}
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("arap::", "")
            key = name.split("<")[0].replace("_kernel", "")
            if key in out:
                continue
            e = {"kernel": name, "report": rep.split("/")[-1]}
            for m, label in WANT.items():
                if m in idx:
                    try:
                        v = float(r[idx[m]].replace(",", ""))
                    except ValueError:
                        continue
                    e[label] = v * SCALE.get(units[idx[m]], 1.0)
            if "dram_read_bytes" in e:
                e["traffic_bytes"] = e["dram_read_bytes"] + e.get("dram_write_bytes", 0.0)
            out[key] = e
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    for k, e in out.items():
        print("%-22s %8.1f us  traffic %8.1f MB  dram %5.1f%%  regs %3d  occ %5.1f%%" % (
            k, e.get("duration_us", 0), e.get("traffic_bytes", 0) / 1e6, e.get("dram_throughput_pct", 0), int(e.get("registers", 0)),
            e.get("achieved_occupancy_pct", 0)))


if __name__ == "__main__":
    main()
