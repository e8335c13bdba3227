#!/bin/bash
# round 2, second sweep: persistent hdiff with dynamic piece scheduling; reference drop-in tests
out=gpurun_out
mkdir -p $out
KB="python -m stencil_benchmarks_b200.scripts.kernel_bench"
python -m pytest tests/test_gpu_dropin.py -x -q 2>&1 | tail -30 > $out/gpu_dropin_r02b.log
for cfg in 0,0,0,0,0,2 0,64,0,0,0,2; do
  echo "== SB200_HDIFF_CFG=$cfg"
  SB200_HDIFF_CFG=$cfg python -m pytest tests/test_gpu_parity.py -x -q -k "hdiff or diffusion" 2>&1 | tail -3
done > $out/hdiff_parity_cfgs_r02b.log 2>&1
timeout 900 $KB --what hdiff --dtypes float64 --repeat 15 --loop 200 --hdiff-sweep \
"0,32,0,0,0,1;0,24,0,0,0,1;0,16,0,0,0,2;0,24,0,0,0,2;0,32,0,0,0,2;0,48,0,0,0,2;0,64,0,0,0,2;0,128,0,0,0,2;0,256,0,0,0,2;0,32,0,0,0,3;0,32,0,0,0,2,2;0,64,0,0,0,2,2;0,128,0,0,0,2,2;0,32,0,1,0,2;0,32,0,3,0,2;0,64,0,3,0,2;0,32,0,4,0,2;0,64,0,1,0,2" \
  > $out/hdiff_persist_sweep_r02b.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
for cfg in 0,32,0,0,0,1 0,32,0,0,0,2 0,64,0,0,0,2; do
  SB200_HDIFF_CFG=$cfg timeout 300 ncu --metrics $M --clock-control none -k regex:hdiff_tma -c 1 --csv \
    --log-file $out/hdiff_ncu_${cfg//,/_}_r02b.csv $KB --what hdiff --dtypes float64 --repeat 1 > /dev/null 2>&1
done
tail -n 40 $out/hdiff_persist_sweep_r02b.log $out/gpu_dropin_r02b.log $out/hdiff_parity_cfgs_r02b.log
