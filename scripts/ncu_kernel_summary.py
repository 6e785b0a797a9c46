"""Summarise one kernel of an .ncu-rep (raw + source pages): headline metrics, opcode mix, stall mix, hottest instructions.\nusage: ncu_kernel_summary.py REPORT [KERNEL_REGEX]"""
import csv, collections, re, subprocess, sys
rep=sys.argv[1]; kern=sys.argv[2] if len(sys.argv)>2 else None
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"]+(["--kernel-name","regex:"+kern] if kern else []),capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[0]; u=rows[1]
want=["Kernel Name","gpu__time_duration.sum","launch__grid_size","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum","l1tex__t_requests_pipe_lsu_mem_global_op_st.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","l1tex__data_pipe_lsu_wavefronts.sum","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers","sm__cycles_elapsed.avg"]
for r in rows[2:]:
    for k,uu,v in zip(h,u,r):
        if k in want: print(f"{k} [{uu}] = {v}")
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"]+(["--kernel-name","regex:"+kern] if kern else []),capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi=next(i for i,r in enumerate(rows) if "Address" in r and "Source" in r)
h=rows[hi]
isrc,ie,iss=h.index("Source"),h.index("Instructions Executed"),h.index("Warp Stall Sampling (All Samples)")
stall_cols=[c for c in h if c.startswith("stall_")]
agg=collections.Counter(); byop=collections.Counter(); st=collections.Counter(); tot=0; top=[]
seen=set()
for r in rows[hi+1:]:
    if r[0] in seen: continue
    seen.add(r[0])
    try: n,s=int(r[ie]),int(r[iss] or 0)
    except: continue
    t=r[isrc].strip()
    op=re.sub(r"^@!?U?P\w+\s+","",t).split()[0] if t else "?"
    byop[op]+=n; st[op]+=s; tot+=n; top.append((s,t))
    for c in stall_cols:
        try: agg[c]+=int(r[h.index(c)] or 0)
        except: pass
print("warp-instr",tot); ts=sum(st.values())
for op,n in st.most_common(14): print(f"  {op:18s} {100*byop[op]/tot:5.1f}% issued {100*n/ts:5.1f}% stall")
for c,n in agg.most_common(8): print(" ",c,f"{100*n/sum(agg.values()):.1f}%")
top.sort(reverse=True)
for s,t in top[:10]: print("  ",s,t)
