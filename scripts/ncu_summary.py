#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here, no GPU needed) into a small text table.
usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [...]"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("launch__registers_per_thread", "regs"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("smsp__issue_active.avg.pct", "issue_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]

def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print("==", rep)
        for r in rows[2:]:
            name = r[idx["Kernel Name"]][:60]
            out = [name]
            for k, short in KEYS:
                if k in idx:
                    v = r[idx[k]].replace(",", "")
                    u = units[idx[k]]
                    try:
                        f = float(v)
                        if short.endswith("_MB"):
                            f = f / {"byte": 1e6, "Kbyte": 1e3, "Mbyte": 1, "Gbyte": 1e-3}.get(u, 1e6)
                        if short == "dur_us":
                            f = f / {"ns": 1e3, "us": 1, "ms": 1e-3, "usecond": 1, "nsecond": 1e3, "msecond": 1e-3}.get(u, 1e3)
                        out.append(f"{short}={f:.1f}")
                    except ValueError:
                        out.append(f"{short}={v}")
            print("  " + " ".join(out))

if __name__ == "__main__":
    main()
