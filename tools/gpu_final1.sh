#!/usr/bin/env bash
# round-end record on ONE GPU: tests, smoke, the default bench line of both arms, the 200-frame C2 record, ncu exports
set -u
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1500 python bench.py > $O/final_default_bar1M_n1.json 2> $O/final_default_bar1M_n1.err ) 2>&1 | grep real
python tools/bench_summary.py $O/final_default_bar1M_n1.json | head -14
( time timeout 1500 python bench.py --impl reference > $O/final_reference_bar1M_n1.json 2> $O/final_reference_bar1M_n1.err ) 2>&1 | grep real
tail -c 600 $O/final_reference_bar1M_n1.json; echo
( time timeout 1500 python bench.py --workload bar17K --steps 200 --warmup 3 --no-secondary > $O/final_bar17K_200frames_n1.json 2> $O/final_bar17K_200frames_n1.err ) 2>&1 | grep real
python tools/bench_summary.py $O/final_bar17K_200frames_n1.json | head -3
bash tools/gpu_ncu_k5.sh bar1M r2f
bash tools/gpu_ncu_k5.sh bar17K r2f
NCU_COUNT=6000 bash tools/gpu_launches.sh bar17K r2f | head -24
