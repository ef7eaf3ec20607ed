set -x
mkdir -p gpurun_out
timeout 900 python bench.py --workload config4 --steps 4 --config4-batches 16,64,256,512 > gpurun_out/c9_config4_1gpu.json 2> gpurun_out/c9_config4.err
echo "config4 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/c9_config4_1gpu.json'))
for r in d['sweep']: print(r['global_batch'], {k: round(v,1) for k,v in r['ms_per_forward'].items()}, r['logits_rel_err_full_vs_reference'])
PY
tail -3 gpurun_out/c9_config4.err
