"""Executed-instruction histogram by opcode for one kernel from `ncu -i X.ncu-rep --page source --csv` output (needs
--import-source on / -lineinfo):  python scripts/ncu_opcode_hist.py source.csv <kernel name substring> [occurrence]"""
import csv,sys,collections,re
# usage: srcsum.py file kernel_substr [occurrence]
f=sys.argv[1]; sub=sys.argv[2]; occ=int(sys.argv[3]) if len(sys.argv)>3 else 0
rows=list(csv.reader(open(f)))
# split into kernels
ks=[]; cur=None
for row in rows:
    if row and row[0]=="Kernel Name":
        cur={"name":row[1],"rows":[]}; ks.append(cur)
    elif cur is not None and row and row[0].startswith("0x"):
        cur["rows"].append(row)
sel=[k for k in ks if sub in k["name"]][occ]
print(sel["name"][:100], len(sel["rows"]),"sass lines")
byop=collections.Counter(); tot=0
for r in sel["rows"]:
    src=r[1].strip(); n=int(r[5] or 0)
    src=re.sub(r'^@!?U?P\d+\s+','',src)
    op=src.split()[0].split('.')[0] if src else '?'
    byop[op]+=n; tot+=n
print("total warp instr",tot)
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
for op,n in byop.most_common(28): print("  %-10s %10d %5.1f%%"%(op,n,100*n/tot))
