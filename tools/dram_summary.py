"""Summarise a single-pass ncu CSV (gpu__time_duration, dram bytes) per process: steady-state averages."""
import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); pi=hdr.index('Process ID')
per=collections.OrderedDict()
for r in rows[1:]:
    d=per.setdefault((r[pi], r[hdr.index('ID')]),{'k':r[ki]})
    d[r[mi]]=float(r[vi].replace(',',''))
procs=collections.OrderedDict()
for (pid,_),d in per.items(): procs.setdefault(pid,[]).append(d)
for pid,ls in procs.items():
    tot=0
    for pat in ('k_link_lane<3','k_node'):
        sel=[d for d in ls if pat in d['k']][-10:]
        if not sel: continue
        avg=lambda m: sum(d[m] for d in sel)/len(sel)
        rd,wr=avg('dram__bytes_read.sum')/1e6,avg('dram__bytes_write.sum')/1e6
        tot+=rd+wr
        print(pid, sel[0]['k'][21:52], 'dur %.1f us'%(avg('gpu__time_duration.sum')/1e3), 'rd %.1f MB wr %.1f MB'%(rd,wr))
    print(pid,'total DRAM per step %.1f MB'%tot)
