"""Summarise an .ncu-rep (raw page + source page) into the handful of numbers we track."""
import csv, subprocess, sys, collections, io

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
    return dict(zip(hdr, vals))

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]

def sass_hist(rep, top=18):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    op = collections.Counter(); thr = collections.Counter(); tot = 0
    for r in rows[2:]:
        if len(r) < 10: continue
        toks = r[ix["Source"]].split()
        o = toks[1] if toks[0].startswith("@") else toks[0]
        o = o.split(".")[0]
        c = int(r[ix["Instructions Executed"]]); op[o] += c; thr[o] += int(r[ix["Thread Instructions Executed"]]); tot += c
    return tot, [(o, 100.0 * c / tot, thr[o] / max(c, 1)) for o, c in op.most_common(top)]

if __name__ == "__main__":
    rep = sys.argv[1]
    r = raw(rep)
    for k in KEYS:
        if k in r: print("%-80s %s" % (k, r[k]))
    tot, hist = sass_hist(rep)
    print("warp instructions executed:", tot)
    for o, pct, t in hist: print("  %-8s %5.1f%%  avg active threads %.1f" % (o, pct, t))
