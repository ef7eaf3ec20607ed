timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; head -c 260 gpurun_out/bench_final.json; echo
