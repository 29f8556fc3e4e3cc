"""Selected metrics of every kernel in an ncu report.   python tools/ncu_metrics.py report.ncu-rep [extra-metric-prefix ...]"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines())); hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_subpipe", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts", "sm__throughput.avg.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__pipe_fma_cycles_active", "sm__inst_executed_pipe_fma",
        "sm__pipe_fmaheavy", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "sm__cycles_elapsed.max", "smsp__inst_executed_pipe_lsu", "sm__pipe_alu_cycles_active",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_lsu", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__pipe_shared_cycles_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__lsu_writeback_active", "l1tex__data_pipe", "smsp__inst_executed_op_shared", "sm__inst_executed_pipe_uniform"] + sys.argv[2:]
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want):
        print(f"{h:88s} {units[i]:10s} {[r[i] for r in rows[2:]]}")
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
try:
    h = rows[1]; idx = {n: i for i, n in enumerate(h)}
    st = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]; tot = {s: 0 for s in st}
    for r in rows[2:]:
        if len(r) < len(h): break
        for s in st:
            try: tot[s] += int(r[idx[s]] or 0)
            except ValueError: pass
    T = sum(tot.values()) or 1
    print("stalls (first kernel):", ", ".join(f"{s[6:]} {100*v/T:.1f}%" for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
except Exception as e:
    print("no source page:", e)
