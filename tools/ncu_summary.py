"""Key metrics of one ncu report (first kernel id).   python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines())); hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_config_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:75s} {rows[1][i]:12s} {[r[i] for r in rows[2:]]}")
# stall breakdown from the source page
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines())); h = rows[1]; idx = {n: i for i, n in enumerate(h)}
st = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]; tot = {s: 0 for s in st}
for r in rows[2:]:
    if len(r) < len(h): break
    for s in st:
        try: tot[s] += int(r[idx[s]] or 0)
        except ValueError: pass
T = sum(tot.values()) or 1
print("stalls:", ", ".join(f"{s[6:]} {100*v/T:.1f}%" for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
