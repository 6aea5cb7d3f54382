"""Prints the headline metrics + top stall reasons of every kernel in an .ncu-rep (via ncu --page raw --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_size" , "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:70])
    for w in want:
        if w in d:
            print("   %-75s %s %s" % (w, d[w], units[hdr.index(w)]))
    stalls = [(float(v.replace(",", "")), k) for k, v in d.items() if "warp_issue_stalled" in k and k.endswith("_per_warp_active.pct") and v not in ("", "n/a")]
    for v, k in sorted(stalls, reverse=True)[:6]:
        print("   stall %-60s %.1f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))
