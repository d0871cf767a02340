#!/bin/bash
# registers / spills per kernel of one translation unit: scripts/ptxas_stats.sh hessian_fast.cu
cd /root/repo/nellie_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC \
  --expt-relaxed-constexpr --expt-extended-lambda -Xptxas=-v -c "$1" -o "${1%.cu}.o" 2>&1 | \
  grep -E "error|warning|Compiling entry|Used|spill" | python3 -c "
import sys,re,subprocess
name=None
for line in sys.stdin:
    if 'error' in line or 'warning' in line: print(line.rstrip()); continue
    m=re.search(r\"Compiling entry function '(\S+)'\",line)
    if m:
        name=subprocess.run(['c++filt',m.group(1)],capture_output=True,text=True).stdout.strip()
        name=re.sub(r'\(anonymous namespace\)::','',name)[:110]; continue
    m=re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores',line)
    if m: spill=m.group(2); continue
    m=re.search(r'Used (\d+) registers',line)
    if m and name: print(f'{m.group(1):>4} regs {spill:>5} spillB  {name}'); name=None
"
