mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02u_bench_N32_T29_8gpu.json 2> gpurun_out/bench8_err.log; echo "bench8 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02u_bench_N32_T29_8gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gather_verified','h2d_ceiling_gbs')}); print('e2e', d['e2e']['value'], 'u8', d.get('e2e_u8',{}).get('value'), 'c2', d['config2'].get('value'), d['config2'].get('gather_verified'), 'sus', d['sustained']['clips_per_s'])
"
