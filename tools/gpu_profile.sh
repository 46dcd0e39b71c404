#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, benches, ncu launch list and full-set captures.  Outputs -> gpurun_out/
# (kept under gpurun's 64 MiB return limit: reports are exported to CSV on the box, big .ncu-rep files are dropped)
set -u
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-r1}
WHAT=${2:-all}
nvidia-smi -L > $O/smi.txt; nproc >> $O/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_bar17K_$TAG.json 2> $O/bench_bar17K_$TAG.err; tail -c 300 $O/bench_bar17K_$TAG.err
timeout 900 python bench.py --workload bar136K_like --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_bar136K_$TAG.json 2> $O/bench_bar136K_$TAG.err; tail -c 300 $O/bench_bar136K_$TAG.err
timeout 900 python bench.py --workload bar1M --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_bar1M_$TAG.json 2> $O/bench_bar1M_$TAG.err; tail -c 300 $O/bench_bar1M_$TAG.err
[ "$WHAT" = "bench" ] && exit 0
# launch list (every launch, device time): set-up + 1 warm-up + 2 timed frames of bar17K
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_bar17K_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-mode > $O/ncu_launch.log 2>&1
export_rep() {  # $1 = report base name
    ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
    ncu -i $O/$1.ncu-rep --page details --csv > $O/$1_details.csv 2>/dev/null
    ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/$1_source.csv.gz
    sz=$(stat -c %s $O/$1.ncu-rep 2>/dev/null || echo 0)
    [ "$sz" -gt 12000000 ] && rm -f $O/$1.ncu-rep
}
# full-set captures (few launches each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_solve_stream' -s 10 -c 2 -f -o $O/prof_solve_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-mode > $O/ncu_solve.log 2>&1
export_rep prof_solve_$TAG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_update|k_trsm|k_potrf|k_extend_add|k_sp_|k_pack' -s 150 -c 14 -f -o $O/prof_factor_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-mode > $O/ncu_factor.log 2>&1
export_rep prof_factor_$TAG
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_energy|k_elem_grad|k_vertex_gather|k_hessian|k_fill|k_solve_stream' -s 8 -c 8 -f -o $O/prof_tet_bar1M_$TAG \
    python bench.py --workload bar1M --steps 1 --warmup 1 --no-cpu-baseline --profile-mode > $O/ncu_tet.log 2>&1
export_rep prof_tet_bar1M_$TAG
du -sh $O; ls -la $O
