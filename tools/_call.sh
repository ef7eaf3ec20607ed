mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "k_extension or fold_downsample or basic_block_chain or golden_frontend" 2>&1 | tail -15
timeout 900 python tools/exp/fold_ds_probe.py > gpurun_out/r02l_fold_ds_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02l_fold_ds_probe.log | tail -30
