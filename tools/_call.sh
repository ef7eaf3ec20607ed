mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "plan or dropout" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/exp/pipeline_switches_probe.py 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r02u_bench_N32_T29.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
python -c "
import json
d=json.load(open('gpurun_out/r02u_bench_N32_T29.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches_per_step')}); print('e2e',d['e2e']['value'], 'u8',d['e2e_u8']['value'], 'lat',d['latency']['ms_per_step'], d['latency']['frac_of_bf16_peak'],'sus', d['sustained']['clips_per_s'], d['sustained']['frac_of_bf16_sustained'], 'c2',d['config2']['value'], d['config2']['latency_ms_unpipelined']); print(d['roofline']['path'])
"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02u_bench_N32_T29_steps20.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r02u_bench_N32_T29_steps20.json')); print('K=20:', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'u8', d['e2e_u8']['value'], 'lat', d['latency']['ms_per_step'])"
