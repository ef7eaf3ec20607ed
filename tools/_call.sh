mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "conv_block or fuse_blocks" 2>&1 | tail -12
timeout 600 python tools/exp/fuse_blocks_probe.py > gpurun_out/r02s_fuse_blocks_probe.log 2>&1; echo rc=$?; tail -22 gpurun_out/r02s_fuse_blocks_probe.log
