out=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555"
timeout 600 $TR bench.py --gpus 8 --steps 200 --warmup 5 --iterate --no-extras --scaling strong > $out/bench_n8_iterate_strong_r02.out 2> $out/bench_n8_iterate_strong_r02.err
timeout 600 $TR bench.py --gpus 8 --steps 200 --warmup 5 --no-extras --scaling strong > $out/bench_n8_strong_r02.out 2> $out/bench_n8_strong_r02.err
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $out/bench_n8_peer_r02.out 2> $out/bench_n8_peer_r02.err
for f in iterate_strong strong peer; do grep "^{" $out/bench_n8_${f}_r02.out | cut -c1-330; tail -n 2 $out/bench_n8_${f}_r02.err; done
