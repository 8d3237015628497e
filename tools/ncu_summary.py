#!/usr/bin/env python
"""Summarise Nsight Compute reports (gpurun_out/*.ncu-rep) into a small JSON / markdown table
for profiles/.  Runs in the GPU-less build container (`ncu -i` only reads the report).

    python tools/ncu_summary.py gpurun_out/prof_big1_fwd.ncu-rep [...] --out profiles/r01_ncu_summary.json
"""
import argparse
import csv
import io
import json
import subprocess
from pathlib import Path

METRICS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "smsp__inst_executed.sum": "instructions",
}


def read_report(path: Path):
    out = subprocess.run(["ncu", "-i", str(path), "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(io.StringIO(out.stdout)))
    header, units = rows[0], rows[1]
    kernels = []
    for values in rows[2:]:
        record = {"report": path.name}
        for name, unit, value in zip(header, units, values):
            if name == "Kernel Name":
                record["kernel"] = value[:120]
            elif name in METRICS or "tensor" in name and "realtime" in name and name.endswith("pct_of_peak_sustained_elapsed"):
                key = METRICS.get(name, name)
                try:
                    record[key] = float(value.replace(",", ""))
                except ValueError:
                    record[key] = value
                record[key + "_unit"] = unit
        kernels.append(record)
    return kernels


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("reports", nargs="+")
    parser.add_argument("--out", required=True)
    args = parser.parse_args()
    summary = []
    for report in args.reports:
        summary.extend(read_report(Path(report)))
    Path(args.out).write_text(json.dumps(summary, indent=1))
    for k in summary:
        print(json.dumps(k))


if __name__ == "__main__":
    main()
