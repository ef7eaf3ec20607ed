N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_r1z_n$N.json 2> gpurun_out/bench_r1z_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r1z_n$N.err; wc -l gpurun_out/bench_r1z_n$N.json; head -c 250 gpurun_out/bench_r1z_n$N.json
