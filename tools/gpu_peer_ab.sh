#!/usr/bin/env bash
# A/B of the peer-memory all-reduce vs ncclAllReduce at N ranks (gpurun --gpus N): parity test, then bench with both
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "two_ranks" 2>&1 | tail -4
for wl in ${WORKLOADS:-bar1M}; do
  for peer in ${PEERS:-1 0}; do
    DOTGPU_PEER_REDUCE=$peer timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $wl --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/peer${peer}_${wl}_n$N.json 2> gpurun_out/peer${peer}_${wl}_n$N.err
    echo "$wl n=$N peer=$peer rc=$?"; python tools/bench_summary.py gpurun_out/peer${peer}_${wl}_n$N.json 2>/dev/null | head -1 || tail -c 800 gpurun_out/peer${peer}_${wl}_n$N.err
  done
done
