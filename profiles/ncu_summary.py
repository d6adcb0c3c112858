"""Key metrics + top stalled instructions of one `ncu --set full --import-source on` report:
    python profiles/ncu_summary.py gpurun_out/<report>.ncu-rep [n_top_instructions]
"""
import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
d=dict(zip(hdr,zip(units,vals)))
keys=['gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__m_xbar2l1tex_read_bytes.sum','launch__grid_size','launch__block_size','sm__cycles_elapsed.avg.per_second','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','dram__bytes_write.sum.per_second','dram__bytes_read.sum.per_second','sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in d: print(k,d[k])
for h in hdr:
    if 'issue_stalled' in h and 'per_issue_active' in h: print('  stall',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), d[h][1])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}; data=rows[2:]
tot=sum(int(r[ix['# Samples']]) for r in data)
print("total samples",tot,"instrs",len(data))
for r in sorted(data,key=lambda r:-int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv)>2 else 20]:
    print("%6s %5.1f%% exec=%8s  %s" % (r[ix['# Samples']], 100*int(r[ix['# Samples']])/tot, r[ix['Instructions Executed']], r[ix['Source']].strip()[:90]))
