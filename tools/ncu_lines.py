#!/usr/bin/env python3
"""Joins an `ncu --page source --csv` SASS dump with `nvdisasm -g -c` line info and prints the
hottest CUDA source lines (samples, instructions, average active threads).
usage: ncu_lines.py <src_sass.csv> <nvdisasm.txt> <file.cu> [top]"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[2:]:
    try:
        sass.append((int(r[col['# Samples']]), int(r[col['Instructions Executed']]), int(r[col['Thread Instructions Executed']]), r[col['Source']].strip(),
                     {k: int(r[col[k]]) for k in col if k.startswith('stall_') and '(' not in k}))
    except Exception:
        pass
# nvdisasm: sequence of instructions with preceding //## File "...", line N
line = None; seq = []
for l in open(sys.argv[2], errors='replace'):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        seq.append(line)
print(len(sass), 'sass rows;', len(seq), 'disasm instrs')
n = min(len(sass), len(seq))
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
tot = 0
for i in range(n):
    s, ins, tins, txt, st = sass[i]
    a = agg[seq[i]]; a[0] += s; a[1] += ins; a[2] += tins; tot += s
    for k, v in st.items(): a[3][k] += v
src = open(sys.argv[3], errors='replace').read().split('\n')
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
print('total samples', tot)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if key is None: continue
    f, ln = key
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    text = src[ln - 1].strip()[:70] if f == sys.argv[3].split('/')[-1] and ln <= len(src) else f
    print(f"{a[0]:7d} {100*a[0]/tot:5.1f}%  inst {a[1]:10d} thr/inst {a[2]/max(a[1],1):5.1f}  L{ln:<4d} {text:70s} {[(k[6:],v) for k,v in st]}")
