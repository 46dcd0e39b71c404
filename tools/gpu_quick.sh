set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_bar17K_$TAG.json 2> $O/bench_bar17K_$TAG.err; tail -c 300 $O/bench_bar17K_$TAG.err
timeout 900 python bench.py --workload bar1M --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_bar1M_$TAG.json 2> $O/bench_bar1M_$TAG.err; tail -c 300 $O/bench_bar1M_$TAG.err
