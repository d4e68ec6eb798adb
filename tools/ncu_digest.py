#!/usr/bin/env python
"""Digest of an .ncu-rep (run where ncu is installed): headline metrics, stall reasons and the hottest
code regions of every captured kernel.  usage: python tools/ncu_digest.py file.ncu-rep [region_size]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__warps_eligible.avg.per_cycle_active"]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    region = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                print("  %-70s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        stalls = []
        for i, h in enumerate(hdr):
            if "smsp__average_warp" in h and "issue_stalled" in h and "not_issued" not in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        print("  stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:8]))
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    starts = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    for k, hi in enumerate(starts):
        hdr = src[hi]
        body = []
        for r in src[hi + 1:]:
            if not r or r[0] == "Kernel Name":
                break
            body.append(r)
        si, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        tot = sum(int(r[si] or 0) for r in body)
        print("-" * 100)
        print("kernel %d: %d SASS instructions, %d stall samples; regions of %d instructions:" % (k, len(body), tot, region))
        for lo in range(0, len(body), region):
            part = body[lo:lo + region]
            s = sum(int(r[si] or 0) for r in part)
            e = sum(int(r[ie] or 0) for r in part)
            if s * 50 < tot:
                continue
            ops = collections.Counter((r[1].split()[1] if r[1].strip().startswith("@") else r[1].split()[0]) for r in part)
            print("  [%5d] samples %6d (%4.1f%%) executed %10d  %s" % (lo, s, 100.0 * s / max(tot, 1), e, ops.most_common(4)))
        top = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:12]
        for i in sorted(top):
            print("    %5d %6s %10s  %s" % (i, body[i][si], body[i][ie], body[i][1][:100]))


if __name__ == "__main__":
    main()
