"""Turn an .ncu-rep (brought back in gpurun_out/) into the small text summary that is committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_kmc_r1_a.ncu-rep profiles/r1_a_kmc_run_kernel.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "inst_executed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "sass__inst_executed_global_loads", "sass__inst_executed_shared_loads",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__average_warps_active_per_inst_executed.ratio",
]
STALL = "smsp__average_warp_latency_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        lines.append("kernel: %s" % d.get("Kernel Name", ("?", ""))[0])
        for k in KEYS:
            if k in d:
                lines.append("  %-72s %s %s" % (k, d[k][0], d[k][1]))
        stalls = sorted(((float(v[0]), h) for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[0] not in ("", "no data")), reverse=True)
        for val, h in stalls[:8]:
            lines.append("  stall %-66s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), val))
        lines.append("")
    text = "\n".join(lines)
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s\n" % rep)
        f.write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
