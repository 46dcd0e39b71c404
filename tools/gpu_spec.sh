#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for sp in 1 0; do
for wl in bar17K bar1M; do
  DOTGPU_SPECULATE=$sp timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/spec${sp}_$wl.json 2> gpurun_out/spec${sp}_$wl.err
  echo "spec=$sp $wl rc=$?"; python tools/bench_summary.py gpurun_out/spec${sp}_$wl.json 2>/dev/null | head -1 || tail -c 800 gpurun_out/spec${sp}_$wl.err
done
done
