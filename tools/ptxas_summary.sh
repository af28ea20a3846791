#!/bin/bash
# registers / spills per kernel of one translation unit (development tool)
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v -c "$1" -o /tmp/ptxas_summary.o 2>&1 | python3 -c "
import sys,re,subprocess
lines=sys.stdin.read().split('\n')
cur=None; spill=''; nerr=0
for l in lines:
    if 'error' in l:
        nerr+=1
        if nerr<8: print(l[:300])
    m=re.search(r'Compiling entry function .(\S+). for',l)
    if m: cur=m.group(1); continue
    if 'Used' in l and cur:
        name=subprocess.run(['c++filt',cur],capture_output=True,text=True).stdout.strip()
        name=re.sub(r'mrl::','',name); name=name.split('(')[0]
        r=re.search(r'Used (\d+) registers',l).group(1)
        print(r, spill, name[:150])
    if 'spill' in l: spill=l.strip().replace('bytes','B').replace(' stack frame','sf').replace(' spill stores','ss').replace(' spill loads','sl')
"
