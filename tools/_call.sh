timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1u.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r1u.log
timeout 300 python bench.py > gpurun_out/bench_r1u.json 2> gpurun_out/bench_r1u.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1u.err; head -c 600 gpurun_out/bench_r1u.json
