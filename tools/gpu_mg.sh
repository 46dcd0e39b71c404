#!/usr/bin/env bash
# multi-GPU loop (gpurun --gpus N): N>1 parity test, then scaling runs of bench.py
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "two_ranks" 2>&1 | tail -8
for wl in ${WORKLOADS:-bar1M}; do
  for n in ${NS:-1 $N}; do
    if [ $n = 1 ]; then
      timeout 900 python bench.py --workload $wl --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/mg_${wl}_n$n.json 2> gpurun_out/mg_${wl}_n$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --workload $wl --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline ${EXTRA:---no-parity --no-secondary} > gpurun_out/mg_${wl}_n$n.json 2> gpurun_out/mg_${wl}_n$n.err
    fi
    echo "$wl n=$n rc=$?"; python tools/bench_summary.py gpurun_out/mg_${wl}_n$n.json 2>/dev/null | head -12 || tail -c 800 gpurun_out/mg_${wl}_n$n.err
  done
done
