"""Turn ncu outputs under gpurun_out/ into the small tracked summaries under profiles/."""
import csv
import subprocess
import sys


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ib, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Block Size"), hdr.index("Grid Size")
    out, tot = [], {}
    for r in rows[1:]:
        name = r[ik].split("(")[0][:90]
        ns = float(r[iv].replace(",", ""))
        out.append((r[0], name, r[ib], r[ig], ns))
        tot[name] = tot.get(name, 0.0) + ns
    total = sum(tot.values())
    with open(dst, "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        fh.write("id,kernel,block,grid,duration_ns\n")
        for r in out:
            fh.write(",".join(str(x).replace(",", ";") for x in r) + "\n")
        fh.write("# share of total device time by kernel\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            fh.write(f"# {100 * v / total:6.2f}%  {v / 1e6:10.3f} ms  {k}\n")


KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "smsp__sass_inst_executed_op_global_st.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
    "smsp__inst_executed_op_branch.sum", "derived__smsp__inst_executed_op_branch_pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    # instruction fetch: requests of the SMs' instruction caches to the GPC-level cache (gcc) and what it forwards to L2
    "gcc__cache_requests_type_instruction.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
    "gcc__average_cache_request_hit_rate.pct", "gcc__xbar2gcc_sectors.sum",
]


def full(rep, dst, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as fh:
        fh.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on ({rep})\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            fh.write(f"\n## {d.get('Kernel Name', '?')}  grid={d.get('Grid Size')} block={d.get('Block Size')}\n")
            u = dict(zip(hdr, units))
            for k in KEYS:
                if k in d:
                    fh.write(f"{k:90s} {d[k]:>22s} {u[k]}\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
