set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02h_bench_N32_T29.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench_N32_T29.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('latency',{}).get('ms_per_step'), d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
print({k: (d[k] if not isinstance(d[k], dict) else {kk: d[k][kk] for kk in list(d[k])[:6]}) for k in d if k in ('sustained','config2')})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02h_launches_forward.csv python tools/prof_target.py forward 2 > gpurun_out/ncu_fwd.log 2>&1; echo "launchlist rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_t -s 1 -c 1 -o gpurun_out/r02h_stem_fused python tools/prof_target.py stemfused 3 > gpurun_out/ncu_stem.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -2
