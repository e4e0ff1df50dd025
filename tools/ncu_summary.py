"""Print selected metrics + top stall lines from an .ncu-rep (needs ncu on PATH). usage: ncu_summary.py rep kernel_regex [nlines]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 25
skip = sys.argv[4] if len(sys.argv) > 4 else "0"   # which match of the regex (0 = first)
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv","-k","regex:"+kre,"-s",skip,"-c","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); hdr=rows[0]; units=rows[1]; r=rows[2]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor']
print('==',r[hdr.index('Kernel Name')][:70])
for w in want:
    if w in hdr: print('  %-62s %s %s'%(w, r[hdr.index(w)], units[hdr.index(w)]))
stall=[h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
st=sorted([(float(r[hdr.index(s)].replace(',','')),s) for s in stall],reverse=True)[:6]
print('   stalls', [(round(v,2), s.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for v,s in st])
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","-k","regex:"+kre,"-s",skip,"-c","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
secs=[(i,rows[i][1], rows[i-1][1] if i>0 else '') for i,x in enumerate(rows) if x and x[0]=='Function Name']
secs.append((len(rows),'',''))
for n,(start,name,path) in enumerate(secs[:-1]):
    end=secs[n+1][0]; h=rows[start+1]
    if '# Samples' not in h: continue
    si=h.index('# Samples'); ii=h.index('Instructions Executed'); lsb=h.index('stall_long_sb')
    data=[];tot=0;ti=0
    for x in rows[start+2:end]:
        if len(x)<=si or not x[0].isdigit(): continue
        try: sm=int(x[si] or 0)
        except: continue
        tot+=sm; ti+=int(x[ii] or 0); data.append((sm,int(x[0]),x[1].strip()[:95],x[ii],x[lsb]))
    if tot==0: continue
    print('#####',path.split('/')[-1],'samples',tot,'warp-instr',ti)
    for d in sorted(data,reverse=True)[:(nl if 'kernels' in path else 6)]: print('  ',d)
