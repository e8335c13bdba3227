#!/bin/bash
# round 2, 4-GPU call: time loop with per-level counters (tests + strong scaling), N=4 weak line, vadv store path
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
python -m pytest tests/test_gpu_multi.py -x -q -k "time_loop or bitwise or partitioned" 2>&1 | tail -15 > $out/multi_gpu_tests_r02_n4.log
timeout 300 python -m stencil_benchmarks_b200.scripts.kernel_bench --what vadv,vadv3 --dtypes float64,float32 --repeat 15 > $out/vadv_storepath_r02.log 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -k "vadv" 2>&1 | tail -3 >> $out/vadv_storepath_r02.log
timeout 600 $TR bench.py --gpus 4 --steps 200 --warmup 5 --iterate --no-extras --scaling strong > $out/bench_n4_iterate_strong_r02.out 2> $out/bench_n4_iterate_strong_r02.err
timeout 600 $TR bench.py --gpus 4 --steps 200 --warmup 5 --no-extras --scaling strong > $out/bench_n4_strong_r02.out 2> $out/bench_n4_strong_r02.err
timeout 900 $TR bench.py --gpus 4 --steps 20 --warmup 5 > $out/bench_n4_peer_r02.out 2> $out/bench_n4_peer_r02.err
timeout 600 $TR bench.py --gpus 4 --steps 20 --warmup 5 --iterate --no-extras > $out/bench_n4_iterate_r02.out 2> $out/bench_n4_iterate_r02.err
tail -n 12 $out/multi_gpu_tests_r02_n4.log; cat $out/vadv_storepath_r02.log
for f in iterate_strong strong peer iterate; do grep "^{" $out/bench_n4_${f}_r02.out | cut -c1-330; tail -n 2 $out/bench_n4_${f}_r02.err; done
