set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
timeout 900 python bench.py --workload config4 --steps 4 > gpurun_out/c12_config4_1gpu.json 2> gpurun_out/c12_config4.err
echo "config4 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/c12_config4_1gpu.json'))
for r in d['sweep']: print(r['global_batch'], {k: round(v,1) for k,v in r['ms_per_forward'].items()}, round(r['b200_full_model_clips_per_s']), round(r['reference_cuda_fp32_clips_per_s']))
PY
tail -3 gpurun_out/c12_config4.err
