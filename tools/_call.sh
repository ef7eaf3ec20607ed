mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "empty_batch or misaligned" 2>&1 | tail -2
SBLK_PROF_STACK=8,2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:encoder_stack -s 1 -c 1 -o gpurun_out/r02k_stack_gpc2 python tools/prof_target.py stack 3 > gpurun_out/ncu_stack.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_stack.log
