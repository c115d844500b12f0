#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers the roofline
discussion in DESIGN.md uses.  Usage: python scripts/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main():
    for path in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        print(f"## {path}")
        for r in rows[2:]:
            print(f"### {r[ix['Kernel Name']]}  (launch id {r[ix['ID']]})")
            for k, label in KEYS:
                if k in ix:
                    print(f"  {label:24s} {r[ix[k]]:>16s} {units[ix[k]]}")
            try:
                rd = float(r[ix['dram__bytes_read.sum']]); wr = float(r[ix['dram__bytes_write.sum']])
                u = units[ix['dram__bytes_read.sum']]
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u]
                t = float(r[ix['gpu__time_duration.sum']])
                tu = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(units[ix['gpu__time_duration.sum']].replace("second", "s"), 1e-6)
                print(f"  {'dram traffic':24s} {(rd + wr) * scale / 1e6:16.1f} MB  -> {(rd + wr) * scale / (t * tu) / 1e9:8.1f} GB/s")
            except Exception as e:  # pragma: no cover
                print("  (traffic n/a)", e)
            print()


if __name__ == "__main__":
    main()
