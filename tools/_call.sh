mkdir -p gpurun_out
timeout 600 python tools/exp/flat2x_modes.py > gpurun_out/r02m_flat2x_modes.log 2>&1; echo "rc=$?"; cat gpurun_out/r02m_flat2x_modes.log
