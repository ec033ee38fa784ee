"""Short text summary of an .ncu-rep (the metrics quoted in DESIGN.md / profiles/README.md): python tools/ncu_summary.py rep [launch index]"""
import csv, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
r = rows[2 + idx]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {rep}, launch {idx}: {r[col['Kernel Name']]}  (ncu --set full --clock-control none)")
for w in want:
    if w in col:
        print(f"{w} = {r[col[w]]} {units[col[w]]}")
print("warp stall reasons, warp-cycles per issued instruction:")
st = [(float(r[i] or 0), h) for h, i in col.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for v, h in sorted(st, reverse=True)[:10]:
    print(f"  {h.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {v:.2f}")
