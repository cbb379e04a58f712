#!/bin/bash
# Dynamic opcode histograms (tools/dyn_by_opcode.py) and per-statement tables (tools/dyn_by_callsite.py) of the flux kernels
# of the CURRENT build, on the bench cases; the .ncu-rep files are processed on the GPU box and deleted (they are ~15 MB each).
#   tools/ncu_opcodes.sh <tag> [case ...]       cases as in tools/prof_case.py, default: the two skin kernels, andreas, ncar
TAG=${1:-r02}; shift
CASES=${@:-"c5:coare3p6+skin c5:ecmwf+skin c5:andreas c5:ncar"}
LIB=aerobulk_b200/libaerobulk_gpu.so
mkdir -p gpurun_out
for c in $CASES; do
  name=$(echo ${c#*:} | sed 's/:/@nb/')
  rep=gpurun_out/op_${TAG}_${name}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_kernel -s 5 -c 1 -f -o $rep \
      python tools/prof_case.py $c > gpurun_out/op_${TAG}_${name}.log 2>&1
  if [ -f $rep.ncu-rep ]; then
    python tools/dyn_by_opcode.py $rep.ncu-rep 40 > gpurun_out/opcodes_${TAG}_${name}.txt 2>&1
    python tools/ncu_summary.py $rep.ncu-rep > gpurun_out/ncu_full_${TAG}_${name}.txt 2>&1
    kern=$(ncu -i $rep.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,subprocess
r=list(csv.reader(sys.stdin)); n=r[2][r[0].index('Kernel Name')]
# mangled name through the demangled one: look it up in the library's symbol table
out=subprocess.run('cuobjdump -elf $LIB | grep -o \"_ZN3abk11flux_kernel[A-Za-z0-9_]*\" | sort -u',shell=True,capture_output=True,text=True).stdout.split()
import re
m=re.search(r'flux_kernel<(\d+), (\d+), (\d+)>',n)
a,s,z=m.groups()
for k in out:
    if k=='_ZN3abk11flux_kernelILi%sELb%sELb%sEEEvNS_8FluxArgsE'%(a,s,z): print(k)
")
    if [ -n "$kern" ]; then
      python tools/dyn_by_callsite.py $rep.ncu-rep $LIB $kern 1 25 > gpurun_out/by_statement_${TAG}_${name}.txt 2>&1
      for op in IMAD.MOV FSEL LDC.64 LOP3 UMOV; do
        python tools/dyn_by_line.py $rep.ncu-rep $LIB $kern "$op" 25 > gpurun_out/byline_${op}_${TAG}_${name}.txt 2>&1
      done
    fi
    rm -f $rep.ncu-rep
  fi
done
ls -la gpurun_out | grep ${TAG}
