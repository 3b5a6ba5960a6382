import csv,re,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
d={h:(v,u) for h,u,v in zip(hdr,units,vals)}
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in d: print(k, d[k])
for h,(v,u) in d.items():
    if re.search(r'smsp__average_warps_issue_stalled.*per_issue_active', h):
        try:
            if float(v)>0.04: print(h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), v)
        except: pass
