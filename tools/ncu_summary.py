"""Key limiter metrics of every kernel in an ncu report (`ncu -i X.ncu-rep --page raw --csv`), one line per launch:
time, DRAM / LSU-data-pipe / issue / ALU / FMA utilisation, resident warps, registers, the top stall reasons.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("us", "gpu__time_duration.sum"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dramR_MB", "dram__bytes_read.sum"),
    ("dramW_MB", "dram__bytes_write.sum"),
    ("lsu_wf%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("warps", "sm__warps_active.avg.per_cycle_active"),
    ("regs", "launch__registers_per_thread"),
    ("inst", "smsp__inst_executed.sum"),
    ("smem_wf", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("lsu_wf", "l1tex__data_pipe_lsu_wavefronts.sum"),
    ("l1_ld_sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
    ("l1_st_sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = d.get("Kernel Name", "?")
            vals = []
            for label, k in KEYS:
                v = d.get(k, "")
                if v in ("", "n/a"):
                    continue
                v = float(v.replace(",", ""))
                if label.endswith("_MB"):
                    unit = u.get(k, "")
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
                if label == "us":
                    unit = u.get(k, "")
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
                vals.append(f"{label}={v:.4g}")
            stalls = sorted(((float(v), k[len(STALL):-len('_per_issue_active.ratio')]) for k, v in d.items()
                             if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
            print(f"{rep.split('/')[-1]} :: {name[:90]}")
            print("   " + " ".join(vals))
            print("   stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))


if __name__ == "__main__":
    main()
