#!/usr/bin/env bash
# grid size of the multi-reduction vector kernels (DOTGPU_MD_BLOCKS): frames/s per workload
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
for wl in ${WORKLOADS:-bar17K bar1M}; do
for b in ${BLOCKS:-296 592 1184}; do
  DOTGPU_MD_BLOCKS=$b timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/md_${wl}_$b.json 2> gpurun_out/md_${wl}_$b.err
  echo "== $wl blocks $b rc=$?"; python tools/bench_summary.py gpurun_out/md_${wl}_$b.json 2>/dev/null | grep -E "fps"
done; done
