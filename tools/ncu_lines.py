#!/usr/bin/env python
"""Per-source-line stall samples / instruction counts of one kernel from an ncu --set full --import-source capture:
the per-instruction CSV of `ncu --page source` joined, instruction by instruction, with the line table of the cubin
(nvdisasm -g on the cubin extracted with cuobjdump -xelf all into /tmp/cub).
usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX MANGLED_NAME_SUBSTRING [TOP_N]"""
import csv, re, subprocess, sys, collections
rep, kern_regex, func_sub = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kern_regex],capture_output=True,text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# may have multiple kernels; take the first block
hdr_i = [i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr = rows[hdr_i]
end = len(rows)
for i in range(hdr_i+1,len(rows)):
    if rows[i] and rows[i][0]=="Kernel Name": end=i;break
ins = rows[hdr_i+1:end]
ci = {k:hdr.index(k) for k in ("Source","# Samples","Instructions Executed","L1 Wavefronts Shared","L1 Wavefronts Shared Excessive")}
dis = subprocess.run(["nvdisasm","-g","/tmp/cub/t2_kernels.sm_100a.cubin"],capture_output=True,text=True).stdout.splitlines()
# locate function
start=None
for i,l in enumerate(dis):
    if l.startswith(".text.") and func_sub in l: start=i;break
lines=[]; cur=None
inl=[]
for l in dis[start+1:]:
    if l.startswith(".text.") or l.startswith("\t.section"): 
        if lines: break
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m2=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m2: lines.append((cur,m2.group(2)))
print("ncu instr",len(ins),"disasm instr",len(lines))
agg=collections.defaultdict(lambda:[0,0,0,0])
n=min(len(ins),len(lines))
tot_s=tot_i=0
for k in range(n):
    r=ins[k]; ln=lines[k][0]
    s=int(r[ci["# Samples"]] or 0); ie=int(r[ci["Instructions Executed"]] or 0)
    wf=int(r[ci["L1 Wavefronts Shared"]] or 0); wx=int(r[ci["L1 Wavefronts Shared Excessive"]] or 0)
    a=agg[ln]; a[0]+=s; a[1]+=ie; a[2]+=wf; a[3]+=wx
    tot_s+=s; tot_i+=ie
src=open("/root/repo/gr-dvbt2ll_b200/csrc/t2_kernels.cu").read().splitlines()
print("total samples",tot_s,"total inst",tot_i)
for ln,a in sorted(agg.items(), key=lambda x:-x[1][0])[:int(sys.argv[4]) if len(sys.argv)>4 else 40]:
    own = ln and ln[0] == "t2_kernels.cu"
    print("%22s samp %5.1f%% inst %5.1f%% smemwf %9d exc %9d | %s" % ("%s:%d" % ln if ln else "?", 100*a[0]/tot_s, 100*a[1]/tot_i, a[2], a[3], src[ln[1]-1].strip()[:100] if own else ""))
