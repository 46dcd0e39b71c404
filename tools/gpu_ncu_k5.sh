#!/usr/bin/env bash
# ncu --set full capture of ONE k_solve_stream launch on a workload -> gpurun_out/prof_k5_<tag>_{raw,details}.csv + source.csv.gz
set -u
mkdir -p gpurun_out; O=gpurun_out; WL=${1:-bar1M}; TAG=${2:-r2}_$WL
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_stream -s 2 -c 1 -f -o $O/prof_k5_$TAG python tools/k5_probe.py $WL 4 > $O/ncu_k5_$TAG.log 2>&1
ncu -i $O/prof_k5_$TAG.ncu-rep --page raw --csv > $O/prof_k5_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_k5_$TAG.ncu-rep --page details --csv > $O/prof_k5_${TAG}_details.csv 2>/dev/null
ncu -i $O/prof_k5_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > $O/prof_k5_${TAG}_source.csv.gz
rm -f $O/prof_k5_$TAG.ncu-rep
tail -2 $O/ncu_k5_$TAG.log | cut -c1-200
