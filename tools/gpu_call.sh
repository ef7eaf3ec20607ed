set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_training_gpu.py -m gpu -q --tb=short -k "matched" 2>&1 | tail -60 > gpurun_out/c4_train_tests.log
cat gpurun_out/c4_train_tests.log
