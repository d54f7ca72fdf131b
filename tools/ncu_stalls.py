#!/usr/bin/env python
"""Summarise an ncu report: per-kernel headline metrics and the source lines with the most stall samples.
usage: tools/ncu_stalls.py report.ncu-rep [kernel-substring] [top-n]"""
import csv
import io
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = raw[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]
    for ki, r in enumerate(raw[2:]):
        name = r[hdr.index("Kernel Name")]
        if sub not in name:
            continue
        print("== launch", ki, name[:60])
        for w in want[1:]:
            if w in hdr:
                print("   %-82s %s %s" % (w, r[hdr.index(w)], raw[1][hdr.index(w)]))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sub]
                                          if sub else ["-i", rep, "--page", "source", "--csv"]))))
    if not src:
        return
    # find header row
    hi = [i for i, r in enumerate(src) if "Source" in r and any("Sampling" in c for c in r)]
    src = [r for r in src if len(r) > 5]
    hi = [i for i, r in enumerate(src) if "Source" in r and any("Sampling" in c for c in r)]
    if not hi:
        print("no source page")
        return
    h = src[hi[0]]
    si = h.index("Source")
    samp = [i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (All")]
    ni = [i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (Not")]
    ii = [i for i, c in enumerate(h) if c == "Instructions Executed"]
    rows = []
    for r in src[hi[0] + 1:]:
        if len(r) <= max(samp + [si]):
            continue
        try:
            rows.append((int(r[samp[0]] or 0), r))
        except ValueError:
            continue
    tot = sum(x for x, _ in rows) or 1
    print("total samples", tot)
    # stall reason columns
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    # totals per stall reason
    agg = {}
    for x, r in rows:
        for i in stall_cols:
            if (r[i] or "0").isdigit():
                agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
    print("stall totals:", " ".join("%s=%.1f%%" % (k.replace("stall_", ""), 100.0 * v / tot)
                                    for k, v in sorted(agg.items(), key=lambda t: -t[1]) if v * 100 > tot))
    print("instructions executed (sum over SASS lines):", sum(int(r[ii[0]] or 0) for _, r in rows if ii and (r[ii[0]] or "0").isdigit()))
    for x, r in sorted(rows, key=lambda t: -t[0])[:top]:
        reasons = sorted(((int(r[i] or 0), h[i]) for i in stall_cols if (r[i] or "0").isdigit()), reverse=True)[:3]
        print("%6.2f%% %7d inst=%-9s %-60s %s" % (100.0 * x / tot, x, r[ii[0]] if ii else "", r[si][:60],
                                               " ".join("%s=%d" % (n.replace("stall_", ""), v) for v, n in reasons if v)))


if __name__ == "__main__":
    main()
