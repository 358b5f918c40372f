import csv,sys,re,subprocess,collections
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
m=dict(zip(hdr,vals))
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','sm__cycles_elapsed.max','smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed','sass__inst_executed_local_loads','sass__inst_executed_local_stores']
for k in keys:
    if k in m: print('%-90s %s %s'%(k,m[k],units[hdr.index(k)]))
for h,v in m.items():
    if 'issue_stalled' in h and 'per_issue_active' in h:
        if float(v)>0.2: print('  stall %-40s %s'%(h.split('stalled_')[1].split('_per_')[0],v))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; body=rows[2:]
ia=hdr.index("Source"); ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples")
tot=sum(int(r[ie]) for r in body); ops=collections.Counter(); samp=collections.Counter()
for r in body:
    op=r[ia].split(); o=op[1] if op[0].startswith('@') else op[0]; o=o.split('.')[0]
    ops[o]+=int(r[ie]); samp[o]+=int(r[isamp])
ts=sum(samp.values())
print('total warp-inst',tot)
for o,c in ops.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 22): print("  %-8s %6.2f%% inst  %6.2f%% samples"%(o,100*c/tot,100*samp[o]/ts))
