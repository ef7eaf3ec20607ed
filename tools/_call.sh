mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_final.json | cut -c1-600
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'u8', d['e2e_u8']['value'], 'lat', d.get('latency',{}).get('ms_per_step'), d['roofline']['kernel'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
print('config2', d['config2']['value'], 'sustained', d['sustained']['clips_per_s'], d['sustained']['frac_of_bf16_sustained'], 'cpu', d['cpu_baseline']['value'])
PY
