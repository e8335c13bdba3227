#!/bin/bash
# round 2, 8-GPU call: the driver's scaling invocation at N=8 (both arms), the time loop, the multi-GPU tests
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $out/bench_n8_peer_r02.out 2> $out/bench_n8_peer_r02.err
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --iterate --no-extras > $out/bench_n8_iterate_r02.out 2> $out/bench_n8_iterate_r02.err
timeout 600 $TR bench.py --gpus 8 --steps 200 --warmup 5 --iterate --no-extras --scaling strong > $out/bench_n8_iterate_strong_r02.out 2> $out/bench_n8_iterate_strong_r02.err
timeout 300 $TR bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $out/bench_reference_n8_r02.out 2> $out/bench_reference_n8_r02.err
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -x -q -k "time_loop or bitwise or partitioned or strong_scaled" 2>&1 | tail -15 > $out/multi_gpu_tests_r02_n8.log
nvidia-smi topo -m > $out/topo_n8.txt 2>&1; nproc >> $out/topo_n8.txt
for f in peer iterate iterate_strong; do grep "^{" $out/bench_n8_${f}_r02.out | cut -c1-400; tail -n 3 $out/bench_n8_${f}_r02.err; done
tail -n 12 $out/multi_gpu_tests_r02_n8.log
