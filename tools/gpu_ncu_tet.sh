#!/usr/bin/env bash
# ncu --set full of the per-tet / fill kernels (K1, K2, K3, K4) on a workload -> gpurun_out/prof_tet_<tag>_{raw,details}.csv
set -u
mkdir -p gpurun_out; O=gpurun_out; WL=${1:-bar1M}; TAG=${2:-r2}_$WL
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_energy|k_grad_block|k_grad_vertex|k_hessian|k_fill" -s ${NCU_SKIP:-9} -c ${NCU_COUNT:-6} -f -o $O/prof_tet_$TAG python tools/k5_probe.py $WL 1 0,1,2,3 > $O/ncu_tet_$TAG.log 2>&1
ncu -i $O/prof_tet_$TAG.ncu-rep --page raw --csv > $O/prof_tet_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_tet_$TAG.ncu-rep --page details --csv > $O/prof_tet_${TAG}_details.csv 2>/dev/null
rm -f $O/prof_tet_$TAG.ncu-rep
tail -2 $O/ncu_tet_$TAG.log | cut -c1-200
python tools/ncu_raw_summary.py $O/prof_tet_${TAG}_raw.csv
