import sys, csv, subprocess
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(raw.splitlines()))
h=r[0]
def col(name): return [i for i,c in enumerate(h) if c==name][0]
names=['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','sm__cycles_elapsed.max','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_shared_ld.sum','sm__inst_executed_pipe_lsu.sum','smsp__cycles_active.avg']
for row in r[2:]:
    for n in names:
        try: print('%-80s %s %s' % (n, row[col(n)][:60], r[1][col(n)]))
        except Exception: pass
    tot=0; st=[]
    for i,c in enumerate(h):
        if 'pcsamp_warps_issue_stalled' in c and 'not_issued' not in c and row[i] not in ('0',''):
            st.append((int(row[i]), c.replace('smsp__pcsamp_warps_issue_stalled_',''))); tot+=int(row[i])
    for v,nme in sorted(st, reverse=True): print('   stall %-22s %6d  %.1f%%' % (nme, v, 100*v/tot))
if len(sys.argv) > 2:
    src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
    r=list(csv.reader(src.splitlines()))
    hi=[i for i,row in enumerate(r) if row and row[0]=='Address'][0]
    h=r[hi]; rows=[x for x in r[hi+1:] if len(x)==len(h)]
    si=h.index('# Samples'); ie=h.index('Instructions Executed'); s_=h.index('Source'); ws=h.index('Warp Stall Sampling (All Samples)')
    tot=sum(int(x[si]) for x in rows if x[si].isdigit())
    print('total samples',tot,'n instr',len(rows))
    top=sorted([(int(x[si]),k) for k,x in enumerate(rows) if x[si].isdigit()], reverse=True)[:int(sys.argv[2])]
    for v,k in sorted(top, key=lambda t:t[1]):
        print('%5d %6d %5.1f%% exec=%s  %s' % (k, v, 100*v/tot, rows[k][ie], rows[k][s_][:90]))
