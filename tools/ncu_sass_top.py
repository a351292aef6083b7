#!/usr/bin/env python
"""The SASS instructions of a kernel that hold the most stall samples (with the three instructions before each):
usage: tools/ncu_sass_top.py REPORT.ncu-rep KERNEL_REGEX TOP_N"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kern],capture_output=True,text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[hi]; ins=[r for r in rows[hi+1:] if len(r)==len(hdr) and r[0]!="Address"]
cs=hdr.index("# Samples"); ce=hdr.index("Instructions Executed")
stall=[c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot=sum(int(r[cs] or 0) for r in ins)
top=sorted(range(len(ins)), key=lambda i:-int(ins[i][cs] or 0))[:int(sys.argv[3])]
for i in top:
    r=ins[i]
    st=sorted([(int(r[hdr.index(c)] or 0),c) for c in stall],reverse=True)[:2]
    print(i, "%.1f%%"%(100*int(r[cs])/tot), r[ce], r[1].strip()[:70], st)
    for j in range(max(0,i-3),i): print("      ", ins[j][1].strip()[:70])
