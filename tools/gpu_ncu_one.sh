#!/bin/bash
# usage: gpu_ncu_one.sh <tag> <kernel regex> <skip> <count> <bench args...>   -> gpurun_out/prof_<tag>_{raw,details}.csv + source.csv.gz
set -u
mkdir -p gpurun_out; O=gpurun_out
TAG=$1; KRE=$2; SKIP=$3; CNT=$4; shift 4
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -f -o $O/prof_$TAG python bench.py "$@" --no-cpu-baseline > $O/ncu_$TAG.log 2>&1
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_$TAG.ncu-rep --page details --csv > $O/prof_${TAG}_details.csv 2>/dev/null
ncu -i $O/prof_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > $O/prof_${TAG}_source.csv.gz
sz=$(stat -c %s $O/prof_$TAG.ncu-rep 2>/dev/null || echo 0); [ "$sz" -gt 12000000 ] && rm -f $O/prof_$TAG.ncu-rep
tail -2 $O/ncu_$TAG.log | cut -c1-300
