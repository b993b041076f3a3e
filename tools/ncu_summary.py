import csv,sys,collections,subprocess
rep=sys.argv[1]; pat=sys.argv[2]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr: print(' ',w.split('.TriageCompute.')[-1],'=',r[hdr.index(w)], rows[1][hdr.index(w)])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+pat],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=None; data=[]
for r in rows:
    if 'Source' in r and 'Instructions Executed' in r: hdr=r; data.append([]); continue
    if hdr is None or len(r)!=len(hdr): continue
    try: n=int(r[hdr.index('Instructions Executed')])
    except: continue
    data[-1].append((r[hdr.index('Source')].strip(),n))
seen=set()
for d in data:
    tot=sum(n for s,n in d)
    if tot in seen: continue
    seen.add(tot)
    segs=[]; cur=None
    for s,n in d:
        if cur is None or n!=cur[0]: cur=[n,[]]; segs.append(cur)
        cur[1].append(s)
    print('== kernel total warp instr',tot,'sass',len(d))
    for n,ss in segs:
        if n*len(ss) < 0.01*tot: continue
        h=collections.Counter((x.split()[1] if x.startswith('@') else x.split()[0]).split('.')[0] for x in ss)
        print('  count=%d n_instr=%d share=%.1f%% %s'%(n,len(ss),100*n*len(ss)/tot,dict(h.most_common(9))))
