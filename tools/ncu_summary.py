#!/usr/bin/env python
"""tools/ncu_summary.py report.ncu-rep [...] -- the handful of ncu raw metrics that back DESIGN.md / bench.py's roofline,
plus the warp-stall sample distribution; plain text for profiles/."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_allocated", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("== %s :: %s" % (rep.split("/")[-1], d.get("Kernel Name", "?")[:110]))
        for k in WANT:
            for h in hdr:
                if h.endswith(k) and (h == k or h.split(".")[-len(k.split(".")):] == k.split(".")):
                    print("   %-82s %s %s" % (k, d[h], u[h]))
                    break
        stalls = sorted(((float(d[h].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in hdr
                         if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and d[h] not in ("", "0")), reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("   stall samples: " + ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in stalls[:7]))
