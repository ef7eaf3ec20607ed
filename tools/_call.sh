mkdir -p gpurun_out
for HB in 3 4 5; do
HEAD_BLOCKS=$HB HEAD_FRACS="0,0.5" VARIANTS=0 timeout 300 python tools/exp/head_frac_probe.py 2>&1 | grep -v "^fuse_prep" | grep pipelined
done > gpurun_out/r02t_head_blocks_probe.log 2>&1
cat gpurun_out/r02t_head_blocks_probe.log
timeout 300 python tools/exp/prep_ahead_probe.py 2>&1 | head -3
