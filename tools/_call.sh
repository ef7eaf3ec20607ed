mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02l_pytest_gpu.log; cat gpurun_out/r02l_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02l_bench_N32_T29.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -3 gpurun_out/bench_err.log
python -c "
import json
d=json.load(open('gpurun_out/r02l_bench_N32_T29.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','gpu_launches_per_step')}); print(d['e2e']['value'], d['e2e_u8']['value'], d['latency']['ms_per_step'], d['sustained']['clips_per_s'], d['config2']['value'])
print(json.dumps(d['roofline'])[:600])
for b in d['breakdown']: print(b)
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02l_launches_forward_N32_T29.csv python tools/prof_target.py forward 2 > gpurun_out/ncu_fwd.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm2 -s 2 -c 2 -o gpurun_out/r02l_fold3 python tools/prof_target.py fold3 3 > gpurun_out/ncu_fold3.log 2>&1; echo "ncu fold3 rc=$?"; tail -2 gpurun_out/ncu_fold3.log
