#!/bin/bash
# usage: gpu_launchlist.sh <tag> <bench args...>  -> gpurun_out/launches_<tag>.csv (+ tests first)
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/launches_$TAG.csv \
    python bench.py "$@" --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1
tail -c 400 $O/ncu_launch_$TAG.log
