#!/bin/bash
# Profiling recipe of one round (run on the GPU box through gpurun); outputs land in gpurun_out/.
#   bash stencil_benchmarks_b200/scripts/profile_round.sh r01
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
KB="python -m stencil_benchmarks_b200.scripts.kernel_bench"

# 1. launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > $out/bench_under_ncu_${tag}.log 2>&1

# 2. full captures of the dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hdiff_tma -c 1 \
    -o $out/hdiff_tma_${tag} $KB --what hdiff --dtypes float64 --repeat 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vadv_onchip -c 1 \
    -o $out/vadv_onchip_${tag} $KB --what vadv --dtypes float64 --repeat 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 1 \
    -o $out/stream_triad_${tag} $KB --what stream --dtypes float64 --repeat 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:basic_kernel -s 14 -c 1 \
    -o $out/basic_lap_${tag} $KB --what basic --dtypes float64 --repeat 1 > /dev/null 2>&1

# 3. STREAM configuration sweep: block,unroll,vector_bytes,streaming
for cfg in 256,1,16,1 256,2,16,1 256,4,16,1 512,2,16,1 512,4,16,1 512,4,16,0 512,8,16,1 1024,4,16,1 \
           256,1,32,1 256,2,32,1 512,1,32,1 512,2,32,1 512,2,32,0 512,4,32,1 1024,2,32,1 128,4,32,1; do
  echo "== SB200_STREAM_CFG=$cfg"
  SB200_STREAM_CFG=$cfg timeout 120 $KB --what stream --dtypes float64 --repeat 10 --stream-log2 28 2>&1 | tail -4
done > $out/stream_sweep_${tag}.log 2>&1

# 4. STREAM size sweep through the plugin class (BASELINE.json configs[1])
timeout 600 python -m stencil_benchmarks_b200.scripts.stream_sweep --max-log2 30 \
    --out $out/stream_sizes_${tag}.csv > $out/stream_sizes_${tag}.log 2>&1

# 5. the reference's own CUDA kernels recompiled for sm_100 (oracle/_ref, built in the dev container)
timeout 600 python -m oracle.ref_cuda --repeat 11 --out $out/reference_cuda_${tag}.json \
    > $out/reference_cuda_${tag}.log 2>&1

# 6. all kernels, both dtypes, for the table in profiles/README.md
timeout 600 $KB --repeat 20 --out $out/kernels_${tag}.json > $out/kernels_${tag}.log 2>&1

# 7. hdiff: rows per march segment, single sweeps and the long loop of bench.py (the two disagree
#    under the board's power cap: profiles/hdiff_segments_r01.log)
{
  for jt in 16 24 32 64 128; do
    echo "== SB200_HDIFF_CFG=0,$jt (single sweeps)"
    SB200_HDIFF_CFG=0,$jt timeout 120 $KB --what hdiff --dtypes float64 --repeat 30 2>&1 | grep hdiff
    echo "== SB200_HDIFF_CFG=0,$jt (bench.py loop)"
    SB200_HDIFF_CFG=0,$jt timeout 300 python bench.py --steps 300 --warmup 10 --no-extras --no-cpu-baseline \
        --e2e-steps 1 | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['clocks'])"
  done
} > $out/hdiff_segments_${tag}.log 2>&1
ls -la $out
