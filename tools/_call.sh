mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "plan or avgpool or golden or dropout" 2>&1 | tail -5
timeout 900 python bench.py --no-config2 --sustained-seconds 0 > gpurun_out/b4.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
python -c "
import json
d=json.load(open('gpurun_out/b4.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches_per_step')}); print('e2e',d['e2e']['value'], 'u8',d['e2e_u8']['value'], 'lat',d['latency']['ms_per_step'], d['latency']['frac_of_bf16_peak'])
for b in d['breakdown']: print(b)
"
