#!/bin/bash
# Profiling recipe of one round (run on the GPU box through gpurun, one GPU); outputs land in gpurun_out/.
#   bash stencil_benchmarks_b200/scripts/profile_round.sh r02
set -u
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
KB="python -m stencil_benchmarks_b200.scripts.kernel_bench"

# 0. the GPU test suite
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/gputests_${tag}.log

# 1. launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras \
    > $out/bench_under_ncu_${tag}.log 2>&1

# 2. full captures of the dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hdiff_tma -c 1 \
    -o $out/hdiff_tma_${tag} $KB --what hdiff --dtypes float64 --repeat 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vadv_onchip -c 1 \
    -o $out/vadv_onchip_${tag} $KB --what vadv --dtypes float64 --repeat 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vadv_onchip -c 1 \
    -o $out/vadv_uvw_${tag} $KB --what vadv3 --dtypes float64 --repeat 1 > /dev/null 2>&1

# 3. all kernels, both dtypes, for the table in profiles/README.md
timeout 900 $KB --what stream,basic,hdiff,vadv,vadv3 --repeat 20 --stream-log2 30 --out $out/kernels_${tag}.json \
    > $out/kernels_${tag}.log 2>&1

# 4. the bench lines themselves (driver's arguments), N=1: both arms, the time loop, vadv
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_n1_${tag}.json 2> $out/bench_n1_${tag}.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_reference_${tag}.json 2> $out/bench_reference_${tag}.err
timeout 600 python bench.py --steps 20 --warmup 5 --iterate --no-extras > $out/bench_n1_iterate_${tag}.json 2> $out/bench_n1_iterate_${tag}.err
timeout 600 python bench.py --steps 200 --warmup 5 --iterate --no-extras > $out/bench_n1_iterate200_${tag}.json 2> $out/bench_n1_iterate200_${tag}.err
timeout 600 python bench.py --workload vadv --steps 20 --warmup 5 --no-extras > $out/bench_n1_vadv_${tag}.json 2> $out/bench_n1_vadv_${tag}.err
ls -la $out | tail -30
tail -n 8 $out/gputests_${tag}.log
