"""Write a compact, committed summary of .ncu-rep captures (read on the CPU box) under profiles/.
usage: python tools/ncu_to_profiles.py <out.md> <title> rep1 [rep2 ...]"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"),
        ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"),
        ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global ld requests"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global st sectors"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait")]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [(dict(zip(hdr, r)), dict(zip(hdr, units))) for r in rows[2:]]


def main():
    out, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = [f"# {title}", "",
             "Source: `ncu --set full --clock-control none --import-source on` on one B200 via gpurun "
             "(cold-cache, serialised launches: compare shares and traffic, not absolute times).", ""]
    for rep in reps:
        for d, u in rows_of(rep):
            lines.append(f"## {rep.split('/')[-1]} — `{d.get('Kernel Name', '?')[:90]}`")
            lines.append("")
            lines.append("| metric | value | unit |")
            lines.append("|---|---:|---|")
            for k, label in KEYS:
                if k in d:
                    lines.append(f"| {label} (`{k}`) | {d[k]} | {u[k]} |")
            lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
