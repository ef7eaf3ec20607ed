mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 50 --warmup 5 --no-config2 --no-u8 --sustained-seconds 0 --no-cpu-baseline $2 > gpurun_out/b6.json 2>/dev/null; python -c "
import json,sys
d=json.load(open('gpurun_out/b6.json')); print(sys.argv[1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('gather_verified'))" "$3"; }
run 29521 "" "gather behind replay"
SBLK_BENCH_GATHER_FIRST=1 run 29522 "" "gather first"
run 29523 "" "gather behind replay"
SBLK_BENCH_GATHER_FIRST=1 run 29524 "" "gather first"
SBLK_BENCH_NO_GATHER=1 run 29525 "" "no gather"
