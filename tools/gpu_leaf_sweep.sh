#!/usr/bin/env bash
# nested-dissection leaf size sweep (DOTGPU_ND_LEAF): frames/s, K5 and factorisation time per workload
set -u
mkdir -p gpurun_out
for wl in ${WORKLOADS:-bar1M}; do
for leaf in ${LEAVES:-12 21 32 48}; do
  DOTGPU_ND_LEAF=$leaf timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/leaf_${wl}_$leaf.json 2> gpurun_out/leaf_${wl}_$leaf.err
  echo "== $wl leaf $leaf rc=$?"; python tools/bench_summary.py gpurun_out/leaf_${wl}_$leaf.json 2>/dev/null | grep -E "fps|factorize|precondition"
done; done
