mkdir -p gpurun_out
timeout 600 python tools/exp/l2_persist_probe.py > gpurun_out/r02q_l2_persist_probe.log 2>&1; echo rc=$?; tail -14 gpurun_out/r02q_l2_persist_probe.log
