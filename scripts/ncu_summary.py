"""Summarise an .ncu-rep: headline metrics + the source lines with the most stall samples."""
import csv, subprocess, sys, io, collections

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:75s} {v} {u}")
for h, u, v in zip(hdr, units, vals):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            if float(v) > 0.3:
                print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:30s} {v}")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[0]
    cols = {n: i for i, n in enumerate(h)}
    samp = next((c for c in h if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"), None)
    srcc = next((c for c in h if c == "Source"), None)
    if samp and srcc:
        agg = collections.Counter()
        for r in rows[1:]:
            try:
                agg[r[cols[srcc]].strip()[:110]] += int(float(r[cols[samp]] or 0))
            except (ValueError, IndexError):
                pass
        tot = sum(agg.values()) or 1
        print(f"-- top source lines by stall samples ({samp}, total {tot})")
        for line, n in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
            print(f"{100*n/tot:5.1f}%  {line}")
    else:
        print("columns:", h[:20])
