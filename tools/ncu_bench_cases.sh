#!/bin/bash
# ncu evidence for the kernels of the CURRENT build on the calls bench.py times (run under gpurun, 1 GPU):
#   tools/ncu_bench_cases.sh <tag>
# per C5 variant: one `--set full` capture (+ the FP64 instruction count) of the 6th flux launch -> summary txt + raw csv;
# C4: andreas / coare3p0 at nb_iter = 30 (per-iteration instruction counts); C2: counters of the 24 launches of a session.
set -u
TAG=${1:-r02}
KEEP_REP=${KEEP_REP:-"ecmwf+skin coare3p6+skin"}
M=sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
mkdir -p gpurun_out
for c in c5:ncar c5:andreas c5:coare3p0 c5:coare3p6 c5:ecmwf c5:coare3p6+skin c5:ecmwf+skin c5:coare3p0+skin c4:andreas:30 c4:coare3p0:30; do
  name=$(echo ${c#*:} | sed 's/:/@nb/')
  rep=gpurun_out/prof_${TAG}_${name}
  timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:flux_kernel -s 5 -c 1 -f -o $rep \
      python tools/prof_case.py $c > gpurun_out/ncu_${TAG}_${name}.log 2>&1
  if [ -f $rep.ncu-rep ]; then
    python tools/ncu_summary.py $rep.ncu-rep > gpurun_out/ncu_full_${TAG}_${name}.txt 2>&1
    ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_${name}_raw.csv 2>/dev/null
    case " $KEEP_REP " in *" $name "*) ;; *) rm -f $rep.ncu-rep ;; esac
  else
    echo "no report for $name" > gpurun_out/ncu_full_${TAG}_${name}.txt
  fi
done
# C2: the 24 launches of the second session (launches 24..47 of flux_kernel)
timeout 900 ncu --metrics gpu__time_duration.sum,$M,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:flux_kernel -s 24 -c 24 --csv --log-file gpurun_out/ncu_metrics_${TAG}_c2_session.csv \
    python tools/prof_case.py c2 > gpurun_out/ncu_${TAG}_c2.log 2>&1
ls -la gpurun_out | grep ${TAG} | tail -40
