timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1z.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r1z.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_r1z.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r1z.log
timeout 300 python bench.py > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1z.err; head -c 300 gpurun_out/bench_r1z.json; echo
