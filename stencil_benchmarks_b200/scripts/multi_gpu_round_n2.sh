#!/bin/bash
# round 2, fourth call (2 GPUs): multi-GPU tests, N=2 bench lines (weak: peer / time loop / NCCL), NVLink counters
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py tests/test_gpu_parity.py -x -q -k "multi or partitioned or time_loop or dropin" 2>&1 | tail -30 > $out/multi_gpu_tests_r02d.log
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2_peer_r02d.json 2> $out/bench_n2_peer_r02d.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --iterate --no-extras > $out/bench_n2_iterate_r02d.json 2> $out/bench_n2_iterate_r02d.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange nccl --no-extras > $out/bench_n2_nccl_r02d.json 2> $out/bench_n2_nccl_r02d.err
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 5 --iterate --no-extras --scaling strong > $out/bench_n2_iterate_strong_r02d.json 2> $out/bench_n2_iterate_strong_r02d.err
timeout 600 ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:hdiff_tma -c 4 --csv --log-file $out/hdiff_peer_nvlink_r02.csv \
  python -c "
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion as h
b = h.Partitioned(domain=(2048, 2048, 80), gpus=2, verify=False, dry_runs=1)
print(b.run())
" > $out/hdiff_peer_nvlink_r02.log 2>&1
nvidia-smi topo -m > $out/topo_n2.txt 2>&1
tail -n 30 $out/multi_gpu_tests_r02d.log; for f in $out/bench_n2_*_r02d.err; do echo $f; tail -n 5 $f; done; tail -n 12 $out/hdiff_peer_nvlink_r02.csv
