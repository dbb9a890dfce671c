import csv, sys, subprocess
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum','launch__registers_per_thread','smsp__inst_executed.sum','sm__icc_request_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum','smsp__inst_executed_op_shared_ld.sum','launch__occupancy_limit','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_fp64.sum']
for h,u,v in zip(hdr,units,vals):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h) or 'launch__occupancy_limit' in h:
        print(f"{h:90s} {u:12s} {v}")
