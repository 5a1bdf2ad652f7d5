import csv, sys, subprocess, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rd=list(csv.reader(raw.splitlines())); hdr=rd[0]; units=rd[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','launch__grid_size','launch__block_size','sm__cycles_elapsed.max']
for row in rd[2:]:
    print('--- kernel', row[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr: print('  %-72s %s %s'%(w,row[hdr.index(w)],units[hdr.index(w)]))
    st=[(float(row[i].replace(',','')),h.replace('smsp__pcsamp_warps_issue_stalled_','')) for i,h in enumerate(hdr) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and row[i] not in ('','n/a')]
    st.sort(reverse=True); print('  stalls:', ', '.join('%s %d'%(n,v) for v,n in st[:8]))
