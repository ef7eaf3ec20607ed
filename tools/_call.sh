mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "plan" 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-config2 --sustained-seconds 0 > gpurun_out/b5.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/b5.json')); print('K=20:', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'u8', d['e2e_u8']['value'], 'lat', d['latency']['ms_per_step'])"
done
