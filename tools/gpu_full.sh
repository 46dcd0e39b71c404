#!/usr/bin/env bash
# full GPU check: all -m gpu tests, then short benches of the two headline workloads (no CPU legs)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for wl in ${WORKLOADS:-bar17K bar1M}; do
  timeout 900 python bench.py --workload $wl --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/full_$wl.json 2> gpurun_out/full_$wl.err
  echo "$wl rc=$?"; python tools/bench_summary.py gpurun_out/full_$wl.json 2>/dev/null | head -12 || tail -c 800 gpurun_out/full_$wl.err
done
