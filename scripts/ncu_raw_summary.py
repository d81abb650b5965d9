"""Key metrics + top stall reasons per kernel from `ncu -i X.ncu-rep --page raw --csv` output:
    python scripts/ncu_raw_summary.py raw.csv"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[0]; idx={k:i for i,k in enumerate(h)}
keys=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","smsp__issue_active.avg.pct","sm__inst_executed.sum","smsp__inst_executed.sum","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active",
"sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
"smsp__thread_inst_executed.sum","smsp__warps_eligible.avg.per_cycle_active","sm__throughput.avg.pct_of_peak_sustained_elapsed","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__inst_executed_op_shared_ld.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stall=[k for k in h if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k]
for r in rows[2:]:
    print("==",r[idx["Kernel Name"]][:70], r[idx.get("Grid Size",0)] if "Grid Size" in idx else "")
    for k in keys:
        if k in idx: print("   %-75s %s %s"%(k,r[idx[k]],rows[1][idx[k]]))
    st=sorted(((float(r[idx[k]] or 0),k) for k in stall),reverse=True)[:7]
    print("   stalls:", ", ".join("%s=%.2f"%(k.split('issue_stalled_')[1].split('_per_')[0],v) for v,k in st))
