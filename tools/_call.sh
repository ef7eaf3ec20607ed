mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm2_block -s 2 -c 1 -o gpurun_out/r02s_block3b python tools/prof_target.py block3b 3 > gpurun_out/ncu_block3b.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_block3b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm2_block -s 2 -c 1 -o gpurun_out/r02s_block3a python tools/prof_target.py block3a 3 > gpurun_out/ncu_block3a.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_block3a.log
