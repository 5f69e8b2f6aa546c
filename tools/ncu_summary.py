import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
H=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','lts__t_bytes.sum','l1tex__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_lsu.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print('-----', r[H.index('Kernel Name')][:70])
    for w in want:
        if w in H:
            i=H.index(w); print(f"  {w:72s} {r[i][:30]:>18s} {units[i]}")
    st=[(float(r[i] or 0),h) for i,h in enumerate(H) if 'smsp__average_warp' in h and 'stall' in h or ('warp_issue_stalled' in h and h.endswith('.pct'))]
    for v,h in sorted(st,reverse=True)[:7]: print(f"     stall {v:8.2f} {h}")
