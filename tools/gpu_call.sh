set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c3_bench_2gpu.json 2> gpurun_out/c3_bench_2gpu.err
echo "bench2 rc=$?"; tail -c 1500 gpurun_out/c3_bench_2gpu.json; tail -3 gpurun_out/c3_bench_2gpu.err
timeout 900 python bench.py --workload config4 --steps 5 > gpurun_out/c3_config4_1gpu.json 2> gpurun_out/c3_config4.err
echo "config4 rc=$?"; tail -c 2500 gpurun_out/c3_config4_1gpu.json; tail -5 gpurun_out/c3_config4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29618 bench.py --workload config4 --gpus 2 --steps 3 --config4-batches 64,512 > gpurun_out/c3_config4_2gpu.json 2> gpurun_out/c3_config4_2gpu.err
echo "config4x2 rc=$?"; tail -c 1200 gpurun_out/c3_config4_2gpu.json; tail -3 gpurun_out/c3_config4_2gpu.err
