import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; vals=rows[-1]
d=dict(zip(hdr,vals))
keys=['launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','launch__occupancy_limit_blocks','launch__shared_mem_config_size','launch__shared_mem_per_block_dynamic','launch__waves_per_multiprocessor','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','smsp__sass_inst_executed_op_shared_ld.sum','smsp__sass_inst_executed_op_shared_st.sum','smsp__sass_inst_executed_op_global_ld.sum','smsp__sass_inst_executed_op_global_st.sum','sm__cycles_elapsed.avg','lts__t_bytes.sum','l1tex__t_bytes.sum']
for k in keys: print(f'{k:80s} {d.get(k)}')
for k,v in d.items():
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
        try:
            if float(v)>0.05: print(f'   stall {k.split("issue_stalled_")[1].split("_per_")[0]:28s} {v}')
        except: pass
for k in ['smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed','smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed','smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed']:
    print(k.split('op_')[1][:5], d.get(k))
