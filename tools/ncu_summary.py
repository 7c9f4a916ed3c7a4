"""Condenses an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers DESIGN.md cites.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [cells_per_launch]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
first = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
rows = rows[first:]
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
]
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name", "?")[:110])
    for k in KEYS:
        if k in d:
            print("  %-62s %16s %s" % (k, d[k], u[k]))
    st = [(h, d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    st = sorted(st, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:8]
    print("  top stall reasons (warps per issue-active cycle): " + ", ".join("%s=%s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for h, v in st))
    if cells:
        inst = float(d["smsp__inst_executed.sum"].replace(",", ""))
        t = float(d["gpu__time_duration.sum"].replace(",", ""))
        tu = u["gpu__time_duration.sum"]
        sec = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu.replace("second", "s").replace("nsecond", "ns"), 1e-3)
        print("  warp-instructions per 32 DP cells: %.2f   (cells per launch %.3g; under ncu %.1f GCUPS, cold/serialised)" % (inst / (cells / 32), cells, cells / sec / 1e9))
