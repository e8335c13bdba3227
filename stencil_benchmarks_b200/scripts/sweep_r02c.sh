#!/bin/bash
# round 2, third call (1 GPU): full GPU test suite, new bench line, reference CUDA sweep with cross-check
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $out/gputests_r02c.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_n1_r02c.json 2> $out/bench_n1_r02c.err
timeout 600 python bench.py --steps 20 --warmup 5 --iterate --no-extras > $out/bench_n1_iterate_r02c.json 2> $out/bench_n1_iterate_r02c.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_reference_r02c.json 2> $out/bench_reference_r02c.err
timeout 1200 python -m oracle.ref_cuda --repeat 7 --check --out $out/reference_cuda_r02.json > $out/reference_cuda_r02.log 2>&1
tail -n 25 $out/gputests_r02c.log; tail -c 1500 $out/bench_n1_r02c.err; tail -c 600 $out/bench_n1_iterate_r02c.err; tail -n 12 $out/reference_cuda_r02.log
