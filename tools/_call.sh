mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -x -q 2>&1 | tail -4
timeout 600 python tools/train_breakdown.py 2>&1 | tail -32 > gpurun_out/r02q_train_step_breakdown_N32_T31.log; head -14 gpurun_out/r02q_train_step_breakdown_N32_T31.log
