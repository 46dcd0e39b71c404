#!/usr/bin/env bash
# launch list (ncu gpu__time_duration.sum, no clock control) of a short bench run -> gpurun_out/launches_<wl>_<tag>.csv.gz + a per-kernel table
set -u
mkdir -p gpurun_out; O=gpurun_out; WL=${1:-bar1M}; TAG=${2:-r2}
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-20000} --csv --log-file $O/launches_${WL}_$TAG.csv \
  python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-secondary --profile-mode > $O/launches_${WL}_$TAG.log 2>&1
python tools/ncu_launches.py $O/launches_${WL}_$TAG.csv ${NCU_SKIP:-0} | tee $O/launches_${WL}_${TAG}_table.txt
gzip -f $O/launches_${WL}_$TAG.csv
