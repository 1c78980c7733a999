#!/usr/bin/env python3
"""Ranks CUDA source lines by warp instructions executed (ncu --page source --csv + nvdisasm -g -c).
usage: ncu_instr.py <src_sass.csv> <nvdisasm.txt> <file.cu> [top] [dump_line]"""
import csv,re,sys
from collections import defaultdict
rows=list(csv.reader(open(sys.argv[1]))); hdr=rows[1]; col={h:i for i,h in enumerate(hdr)}
sass=[]
for r in rows[2:]:
    try: sass.append((int(r[col['# Samples']]),int(r[col['Instructions Executed']]),int(r[col['Thread Instructions Executed']]),r[col['Source']].strip()))
    except Exception: pass
line=None; seq=[]
for l in open(sys.argv[2],errors='replace'):
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: line=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S',l): seq.append(line)
tot=sum(s[1] for s in sass); print('total warp instr',tot,'samples',sum(s[0] for s in sass))
agg=defaultdict(lambda:[0,0,0])
for i,s in enumerate(sass):
    a=agg[seq[i]]; a[0]+=s[0]; a[1]+=s[1]; a[2]+=s[2]
src=open(sys.argv[3]).read().split('\n'); name=sys.argv[3].split('/')[-1]
top=int(sys.argv[4]) if len(sys.argv)>4 else 25
for k,a in sorted(agg.items(),key=lambda kv:-kv[1][1])[:top]:
    f,ln=k if k else ('?',0)
    print(f"{a[1]:11d} {100*a[1]/tot:5.1f}% samp {a[0]:6d} thr {a[2]/max(a[1],1):4.1f} {f}:{ln} {src[ln-1].strip()[:80] if f==name else ''}")
if len(sys.argv)>5:
    for i,s in enumerate(sass):
        if seq[i]==(name,int(sys.argv[5])): print(s)
