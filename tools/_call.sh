mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu_s2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_s2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_s2.log
timeout 600 python bench.py > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('latency',{}).get('ms_per_step'), d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['roofline']['path'])
print({k: d[k] for k in d if k in ('sustained','config2','gpu_launches','clocks')})
PY
