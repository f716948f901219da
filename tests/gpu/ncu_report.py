"""Summarise an ncu --set full report (.ncu-rep) read on the CPU box: one dict per captured launch."""
import csv
import subprocess
import sys

WANT = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'ms'), ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('l1tex__t_sector_hit_rate.pct', 'l1hit%'), ('lts__t_sector_hit_rate.pct', 'l2hit%'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%'), ('smsp__inst_executed.sum', 'inst'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'st_long'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'st_short'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'st_wait'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'st_bar'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'st_math'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'st_notsel'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'st_lg'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'st_mio'),
        ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'st_br'),
        ('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'st_disp')]


def summarise(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for w, n in WANT:
            if w in hdr:
                v = r[hdr.index(w)]
                try:
                    v = round(float(v), 3)
                except ValueError:
                    v = v[:48]
                u = units[hdr.index(w)]
                d[n + ("[" + u + "]" if n in ("ms", "rd", "wr") else "")] = v
        res.append(d)
    return res


if __name__ == "__main__":
    for d in summarise(sys.argv[1]):
        print(d)
