"""Summarise an .ncu-rep (read here, no GPU needed): one block per profiled launch with the metrics the
roofline discussion uses.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
    "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_set_accesses_pipe_tex_mem_texture.sum", "sm__inst_executed_pipe_tex.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} profiled launches (ncu --set full --clock-control none; replayed, cold cache)")
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']]}  [id {r[col['ID']]}]")
        for w in WANT:
            if w in col and r[col[w]] != "":
                print(f"  {w:85s} {r[col[w]]:>16s} {units[col[w]]}")


if __name__ == "__main__":
    main()
