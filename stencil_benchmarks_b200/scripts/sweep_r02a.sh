#!/bin/bash
# round 2, first kernel sweep: persistent hdiff schedules, merged vadv
out=gpurun_out
mkdir -p $out
KB="python -m stencil_benchmarks_b200.scripts.kernel_bench"
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $out/gputests_r02a.log
for cfg in 0,0,0,0,0,3 0,0,0,0,0,1 0,64,0,0,0,2 0,0,0,3,0,2; do
  echo "== SB200_HDIFF_CFG=$cfg"
  SB200_HDIFF_CFG=$cfg python -m pytest tests/test_gpu_parity.py -x -q -k "hdiff or diffusion" 2>&1 | tail -3
done > $out/hdiff_parity_cfgs_r02a.log 2>&1
timeout 900 $KB --what hdiff --dtypes float64 --repeat 15 --loop 200 --hdiff-sweep \
"0,32,0,0,0,1;0,16,0,0,0,1;0,16,0,0,0,2;0,24,0,0,0,2;0,32,0,0,0,2;0,48,0,0,0,2;0,64,0,0,0,2;0,128,0,0,0,2;0,256,0,0,0,2;0,512,0,0,0,2;0,0,0,0,0,3;0,32,0,0,0,2,2;0,128,0,0,0,2,2;0,0,0,0,0,3,2;0,32,0,1,0,2;0,32,0,2,0,2;0,32,0,3,0,2;0,32,0,4,0,2;0,128,0,3,0,2;0,0,0,3,0,3;0,0,0,1,0,3" \
  > $out/hdiff_persist_sweep_r02a.log 2>&1
timeout 600 $KB --what vadv,vadv3 --dtypes float64,float32 --repeat 15 > $out/vadv_merged_r02a.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:vadv_onchip -c 2 --csv --log-file $out/vadv_merged_dram_r02a.csv \
  $KB --what vadv3 --dtypes float64 --repeat 1 > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_n1_r02a.json 2> $out/bench_n1_r02a.err
tail -n 40 $out/hdiff_persist_sweep_r02a.log $out/vadv_merged_r02a.log $out/gputests_r02a.log $out/hdiff_parity_cfgs_r02a.log
