#!/bin/bash
# One `ncu --set full` capture per flux-kernel variant of the CURRENT build (run under gpurun, 1 GPU):
#   tools/ncu_all.sh <tag> [Ni Nj]
# For each (algo, skin, day/night) case tools/kbench.py runs a 12-step device-resident session; the 6th flux launch is
# captured with source counters.  Summaries (tools/ncu_summary.py) go to gpurun_out/ncu_full_<tag>_<case>.txt, the raw
# metric page to ..._raw.csv; the .ncu-rep files are kept only for the cases named in KEEP_REP (64 MiB pull limit).
set -u
TAG=${1:-r02}
NI=${2:-1440}
NJ=${3:-720}
KEEP_REP=${KEEP_REP:-"coare3p6_skin_day andreas"}
mkdir -p gpurun_out
for c in "ncar,0,-" "andreas,0,-" "coare3p0,0,-" "coare3p6,0,-" "ecmwf,0,-" "coare3p6,1,night" "coare3p6,1,day" "ecmwf,1,night" "ecmwf,1,day" "coare3p0,1,day" "andreas,0,-,30" "coare3p0,0,-,30"; do
  IFS=, read algo skin rad nb <<< "$c"
  name=${algo}$([ "$skin" = 1 ] && echo "_skin_${rad}")$([ -n "$nb" ] && echo "_nb${nb}")
  export KBENCH_NITER=${nb:-5}
  c="$algo,$skin,$rad"
  rep=gpurun_out/prof_${TAG}_${name}
  KBENCH_ONE=$c timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_kernel -s 5 -c 1 -f -o $rep \
      python tools/kbench.py $NI $NJ > gpurun_out/ncu_${TAG}_${name}.log 2>&1
  if [ -f $rep.ncu-rep ]; then
    python tools/ncu_summary.py $rep.ncu-rep > gpurun_out/ncu_full_${TAG}_${name}.txt 2>&1
    ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_${name}_raw.csv 2>/dev/null
    case " $KEEP_REP " in *" $name "*) ;; *) rm -f $rep.ncu-rep ;; esac
  else
    echo "no report for $name" > gpurun_out/ncu_full_${TAG}_${name}.txt
  fi
done
ls -la gpurun_out | tail -40
