"""Key metrics of an .ncu-rep (raw page) as JSON: python tools/ncu_summary.py rep [out.json]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__cycles_active.avg','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_fp64.sum','lts__t_bytes.sum','l1tex__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.per_cycle_active','launch__occupancy_limit_blocks','launch__waves_per_multiprocessor','sm__maximum_warps_per_active_cycle_pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed']
out = {}
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and h.endswith('.ratio')) or 'pipe_fp64' in h:
        out[h] = "%s %s" % (vals[i], units[i])
s = json.dumps(out, indent=1)
print(s)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(s)
