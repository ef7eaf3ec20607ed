set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_training_gpu.py -m gpu -q --tb=short 2>&1 | tail -15
python tools/train_breakdown.py 32 > gpurun_out/c8_train_breakdown.log 2>&1; head -16 gpurun_out/c8_train_breakdown.log
timeout 900 python bench.py --workload config3 --steps 5 > gpurun_out/c8_config3_b256_1gpu.json 2> gpurun_out/c8_config3.err
echo "config3 rc=$?"; cat gpurun_out/c8_config3_b256_1gpu.json; tail -3 gpurun_out/c8_config3.err
