import csv, collections, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__cycles_active.avg','sm__cycles_elapsed.max','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','sm__cycles_active.avg']
for i,h in enumerate(hdr):
    if h in want: print('%-70s %-12s %s'%(h,units[i],vals[i]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
ops=collections.Counter(); tot=0; stall=collections.Counter(); samples=0
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
per=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    try: n=int(r[ix['Instructions Executed']])
    except: continue
    s=r[ix['Source']].strip(); op=s.split()[0] if s else '?'
    if op.startswith('@'): op=s.split()[1]
    op=op.split('.')[0]; ops[op]+=n; tot+=n
    sm=int(r[ix['# Samples']] or 0); samples+=sm
    for c in stall_cols:
        try: stall[c]+=int(r[ix[c]])
        except: pass
    per.append((sm,n,r[ix['Address']][-5:],s[:80],{c:int(r[ix[c]] or 0) for c in stall_cols if (r[ix[c]] or '0')!='0'}))
print('total warp instr',tot)
for k,v in ops.most_common(14): print('  %-10s %12d %5.1f%%'%(k,v,100*v/tot))
print('samples',samples)
for k,v in stall.most_common(10): print('  %-26s %8d %5.1f%%'%(k,v,100*v/max(1,samples)))
print('top sampled instructions:')
for sm,n,a,s,d in sorted(per,reverse=True)[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print('  %6d %9d %s %-70s %s'%(sm,n,a,s,{k.replace("stall_",""):v for k,v in sorted(d.items(),key=lambda kv:-kv[1])[:3]}))
